timeout 600 python -m pytest tests/test_gpu_seq2seq.py -x -q -m gpu 2>&1 | tail -6
timeout 200 python tools/seq2seq_bench.py 2>&1 | tail -1
timeout 200 python tools/seq2seq_bench.py --tf32 2>&1 | tail -1
