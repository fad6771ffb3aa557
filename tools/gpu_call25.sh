cd $GRAFT_REPO_ROOT
bash tools/run_sanitizer.sh r02e
cuobjdump -sass obs-color-monitor_b200/lib/libscope_b200.so | grep "Function :" | sed "s/^\s*//" > gpurun_out/sanitizer/functions_r02e.txt
bash tools/run_ncu.sh r02e 64
