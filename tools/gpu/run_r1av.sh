timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k short_templates 2>&1 | grep -E "^E|assert|Error" | head -30
JB_NO_ROOT_FUSION=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k short_templates 2>&1 | tail -3
