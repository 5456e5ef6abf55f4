export NTB200_BF16X3_MIN_ROWS=1
timeout 120 python tools/gemm_check.py all 1 128 64 128
timeout 120 python tools/gemm_check.py all 1 256 128 96
timeout 120 python tools/gemm_check.py all 1 1000 384 1152
timeout 120 python tools/gemm_check.py all 1 4900 96 288
timeout 120 python tools/gemm_check.py all 1 777 200 264
timeout 120 python tools/gemm_check.py all 1 16384 384 1536
