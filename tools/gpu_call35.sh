cd $GRAFT_REPO_ROOT
bash tools/gpu_final.sh final3 r02h
bash tools/run_sanitizer.sh r02h
cuobjdump -sass obs-color-monitor_b200/lib/libscope_b200.so | grep "Function :" | sed "s/^\s*//" > gpurun_out/sanitizer/functions_r02h.txt
