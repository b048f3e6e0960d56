O=gpurun_out/r02san; mkdir -p $O
for tool in memcheck racecheck synccheck; do
DRL_B200_CUDA_GRAPH=0 timeout 600 compute-sanitizer --tool $tool --log-file $O/sanitizer_final_$tool.log python tools/sanitize_step.py --modes 5 --batch 4 --trajectory 20 --steps 2 > $O/sanitizer_final_$tool.out 2>&1
echo "$tool rc=$?"; tail -n 1 $O/sanitizer_final_$tool.out; tail -n 2 $O/sanitizer_final_$tool.log
done
