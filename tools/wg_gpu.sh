timeout 200 python tools/wg_dbg.py 2>&1 | grep -E "weights|ratio" | head -12
timeout 300 python -m pytest tests/test_gpu_train.py -x -q -m gpu 2>&1 | tail -4
timeout 120 python tools/train_prof.py 50 2>&1 | tail -1
