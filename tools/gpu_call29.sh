cd $GRAFT_REPO_ROOT
bash tools/gpu_final.sh final2 r02f
bash tools/run_sanitizer.sh r02f
cuobjdump -sass obs-color-monitor_b200/lib/libscope_b200.so | grep "Function :" | sed "s/^\s*//" > gpurun_out/sanitizer/functions_r02f.txt
