for d in 3 2 1 0; do echo "== CV_FE3_DEBUG=$d"; CV_FE3_DEBUG=$d timeout 60 python tools/gpu_slots.py 64 2>&1 | tail -2 | cut -c1-200; done
