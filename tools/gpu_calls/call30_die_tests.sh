# the two die-walk tests after the split (bit identity in-process, calibration verdict from a fresh process)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -rs -k "die" 2>&1 | tail -5
