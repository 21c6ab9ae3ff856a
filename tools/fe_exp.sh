timeout 600 python -m pytest tests -m gpu -q -s -x -k "fused or every_layer_matches_oracle_bf16 or forward_bf16" 2>&1 | grep -E "rel err|passed|failed|Error|error|assert" | tail -12
echo "$(timeout 120 python tools/gpu_slots.py 4096 2>&1 | tail -5 | cut -c1-200)"
