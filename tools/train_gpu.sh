timeout 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_multirank.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/train_prof.py 50 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --launch-skip 150 -c 120 --csv --log-file gpurun_out/train_launches.csv python tools/train_prof.py 6 > gpurun_out/train_ncu.log 2>&1
python tools/train_launches.py gpurun_out/train_launches.csv
