# Build here first: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/ubench2 tools/ubench2.cu
# (the binary travels with the snapshot).  Under gpurun: about a minute of GPU time.
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.current.sm,clocks.max.sm --format=csv > gpurun_out/ubench2.txt
timeout -s KILL 240 ./tools/ubench2 >> gpurun_out/ubench2.txt 2>&1
tail -5 gpurun_out/ubench2.txt
