timeout 600 python -m pytest tests -m gpu -q -s -x -k "fused_front or other_board or forward_bf16" 2>&1 | grep -E "front end|passed|failed|rror|assert" | tail -8
echo "$(timeout 120 python tools/gpu_slots.py 4096 2>&1 | tail -1 | cut -c1-200)"
