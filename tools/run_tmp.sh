python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
bash tools/ab.sh 0 stock pf6 2>&1 | grep -E "==|call|preprocess"
