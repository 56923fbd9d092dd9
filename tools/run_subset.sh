cd /root/repo 2>/dev/null || cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_parity.py tests/test_c_abi.py -m gpu -q -k "free_drift or c_caller or configuration_switches or bench_configuration or immersed or coastline" 2>&1 | tail -30
