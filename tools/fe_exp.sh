timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "$(timeout 120 python tools/gpu_slots.py 4096 2>&1 | tail -1 | cut -c1-200)"
