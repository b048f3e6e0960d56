export DRL_B200_DGRAD_GATHER=1
bash tools/gpu_round2_f.sh
