timeout 900 ncu --set full --clock-control none --import-source on -k regex:tree_bwd --launch-skip 60 -c 6 -o gpurun_out/bwd_full python tools/train_prof.py 6 > gpurun_out/bwd_full.log 2>&1
tail -3 gpurun_out/bwd_full.log
