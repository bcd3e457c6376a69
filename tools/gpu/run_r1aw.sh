timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k short_templates 2>&1 | grep -E "^E|passed|failed" | head -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/gpu/sharded_check.py 2>&1 | grep -E "identical|Error|error" | head
