timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for i in 1 2; do
timeout 200 python tools/seq2seq_bench.py 2>&1 | tail -1
N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_s2s6.so timeout 200 python tools/seq2seq_bench.py 2>&1 | tail -1
done
N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_s2s6.so timeout 600 python -m pytest tests/test_gpu_seq2seq.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_s2s.json 2> gpurun_out/bench_s2s.err; tail -c 600 gpurun_out/bench_s2s.err
