cd $GRAFT_REPO_ROOT
O=gpurun_out/final2; mkdir -p $O
timeout -s KILL 200 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -4 $O/pytest.full > $O/pytest.log
( SCOPE_TMA_X0=1 SCOPE_TMA_L2=0 timeout -s KILL 60 python -m pytest -x -q -m gpu tests/test_gpu_parity.py -k unaligned 2>&1 | tail -3 ) > $O/x0_l2none.log 2>&1
( SCOPE_TMA_X0=1 timeout -s KILL 60 compute-sanitizer --tool memcheck python -m pytest -x -q -m gpu tests/test_gpu_parity.py -k unaligned 2>&1 | grep -v "^$" | head -40 ) > $O/x0_sanitizer.log 2>&1
echo "== pytest"; cat $O/pytest.log; echo "== x0 with L2 promotion none"; cat $O/x0_l2none.log; echo "== sanitizer"; head -30 $O/x0_sanitizer.log
