O=gpurun_out/r02v; mkdir -p $O
F=tools/microbench/bulk_gemm_feed
{
$F 5 27 16 32 32 3 1 1 1 0      # lstm_dgrad as is
$F 5 27 16 32 32 3 0 1 1 0      # private tiles
$F 5 27 16 32 32 3 1 2 1 0      # two issuing lanes
$F 5 27 16 32 32 3 1 1 4 0      # 4 chunks per operand
$F 5 27 16 32 32 3 1 1 1 400    # consumer holds the stage 0.4 us (MMA time)
$F 5 27 32 16 16 6 1 1 1 0      # half-size stages, 6 deep
$F 5 4 9 32 64 2 1 1 1 0        # lstm_fwd split (20 CTAs per split; 7 splits -> use 5x28)
$F 5 28 9 32 64 2 1 1 1 0       # lstm_fwd: 140 CTAs
$F 29 4 9 32 64 2 1 1 1 0       # lstm_wgrad: 116 CTAs
$F 29 4 9 32 64 2 1 2 1 0
$F 29 4 18 16 32 4 1 2 1 0      # half-size stages, 4 deep
} > $O/bulk_gemm_feed.txt 2>&1
cat $O/bulk_gemm_feed.txt
