timeout 1200 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_config3_full_size_identity_and_bounds 2>&1 | tail -4
timeout 300 python tools/api_latency.py 100 1000 2000 3000 4000 8192 2>&1 | tail -7
