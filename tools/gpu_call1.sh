cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/call1_gpu.txt; nproc >> gpurun_out/call1_gpu.txt
bash tools/run_ubench2.sh
bash tools/run_ab.sh quick wide_straight_dephase wide_dephase dephase wide_straight wide straight immcoef w8_straight w8 w8_dephase r120 nopipe ballot rawflat deepring 2>&1 | tail -60
bash tools/run_sanitizer.sh r02a
