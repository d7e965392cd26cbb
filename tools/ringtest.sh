L=$PWD/phfpfac_b200/_build/libpfac_b200_nos1.so
for st in 4 6 9; do echo "== nos1 stages $st"; PFAC_RING_STAGES=$st PFAC_B200_LIB=$L python tools/microbench.py --sizes 1024 --iters 5 2>&1 | grep -E "GB/s" | sed -E "s/info=.*//"; done
echo "== nos1 T3=8K"; PFAC_T3_BYTES=8192 PFAC_B200_LIB=$L python tools/microbench.py --sizes 1024 --iters 5 2>&1 | grep -E "GB/s|derived" | sed -E "s/info=.*//; s/.*(ring_stages.: [0-9]+).*/\1/"
echo "== nos1 T3=8K TM2=0"; PFAC_T3_BYTES=8192 PFAC_TM2_BYTES=0 PFAC_B200_LIB=$L python tools/microbench.py --sizes 1024 --iters 5 2>&1 | grep -E "GB/s|derived" | sed -E "s/info=.*//; s/.*(ring_stages.: [0-9]+).*/\1/"
L=$PWD/phfpfac_b200/_build/libpfac_b200.so
for st in 4 6 9; do echo "== base stages $st"; PFAC_RING_STAGES=$st PFAC_B200_LIB=$L python tools/microbench.py --sizes 1024 --iters 5 2>&1 | grep -E "GB/s" | sed -E "s/info=.*//"; done
echo "== base T3=8K"; PFAC_T3_BYTES=8192 PFAC_B200_LIB=$L python tools/microbench.py --sizes 1024 --iters 5 2>&1 | grep -E "GB/s|derived" | sed -E "s/info=.*//; s/.*(ring_stages.: [0-9]+).*/\1/"
