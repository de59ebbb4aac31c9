python -m pytest tests/test_gpu_parity.py tests/test_gpu_widening.py tests/test_plugins.py -m gpu -q --tb=short -x 2>&1 | tail -3
