timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
bash tools/train_gpu.sh 2>&1 | grep -E "train step|adam|norm|repack|conv_quad|sum "
