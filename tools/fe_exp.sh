timeout 600 python -m pytest tests -m gpu -q -s -x -k "fused or forward_bf16" 2>&1 | grep -E "front end|passed|failed|Error|error|assert" | tail -8
echo "$(timeout 120 python tools/gpu_slots.py 4096 2>&1 | tail -5 | cut -c1-200)"
