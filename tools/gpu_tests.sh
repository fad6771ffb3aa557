cd $GRAFT_REPO_ROOT
O=gpurun_out/${1:-t}; mkdir -p $O
timeout -s KILL 1200 python -m pytest -x -q -m gpu tests > $O/pytest.full 2>&1; echo "exit $?" >> $O/pytest.full; tail -25 $O/pytest.full
