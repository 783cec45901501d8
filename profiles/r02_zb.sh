# round 2, call zb: tiled block-per-env kernel, 4 UEs per thread / 3 CTAs per SM against 2 UEs per thread / 4 CTAs
set -x
timeout 900 python profiles/variant_sweep.py run "mobile-synthetic-central-v0:16384,mobile-synthetic-ma-v0:8192" 512 > gpurun_out/r02_zb_variants.txt 2>&1
cat gpurun_out/r02_zb_variants.txt
MBE_LIB_PATH=$PWD/mobile_env_gan_b200/csrc/libmbe_BIG_FORCE_MAXI40.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_more.py -m gpu -q --maxfail=5 -k "wide or synthetic or big or pf" 2>&1 | tail -3
