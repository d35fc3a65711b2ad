timeout 600 python -m pytest tests/test_box_gpu.py -x -q 2>&1 | tail -4
timeout 300 python tools/hbm_kernels_gpu.py 2>&1 | grep -i "box\|torch"
timeout 600 python tools/variant_sweep.py
