timeout 600 python -m pytest tests -m gpu -q -x -k "fused or forward_bf16 or every_layer or split" 2>&1 | tail -2
echo "$(timeout 120 python tools/gpu_slots.py 4096 2>&1 | tail -1 | cut -c1-200)"
