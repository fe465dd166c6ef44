cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2; do timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_r3d_$i.log 2>&1; echo "pytest $i exit=$?"; tail -n 2 gpurun_out/pytest_gpu_r3d_$i.log; grep FAILED gpurun_out/pytest_gpu_r3d_$i.log; done
