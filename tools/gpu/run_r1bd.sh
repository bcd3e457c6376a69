(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "overflow or workspace") 2>&1 | grep -E "^E|passed|failed|Error" | head -20
