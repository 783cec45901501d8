# round 2, call zm: the GPU suite on the fallback kernel families (UEs-per-thread and thread-per-env kernels off; PDL off)
set -x
MBE_UPT=0 MBE_TPE=0 MBE_PDL=0 timeout 1800 python -m pytest tests -m gpu -q --maxfail=30 2>&1 | tail -25 | tee gpurun_out/r02_zm_pytest_fallback.txt
