python tools/gemm_bench.py 24000 2304 768 --iters 20 --counters
python tools/gemm_bench.py 24000 6144 768 --iters 20 --counters
python tools/gemm_bench.py 24000 768 3072 --iters 20 --f32out --resid --counters
python tools/gemm_bench.py 24000 768 3072 --iters 20 --counters
python tools/gemm_bench.py 24000 768 768 --iters 20 --f32out --resid --counters
python tools/gemm_bench.py 384000 2304 768 --iters 5 --counters
python tools/gemm_bench.py 384000 768 3072 --iters 5 --f32out --resid --counters
