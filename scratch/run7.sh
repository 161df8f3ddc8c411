python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor_core" 2>&1 | tail -40 | cut -c1-300
python -m pytest tests -m gpu -q 2>&1 | tail -8
python tools/_sweep.py 16 2>&1 | tail -1
