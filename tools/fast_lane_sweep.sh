for fw in 4 2 1; do for th in 0.7 0.4; do
echo "== RR_FAST_WARPS=$fw RR_FAST_THRESHOLD=$th"
RR_FAST_WARPS=$fw RR_FAST_THRESHOLD=$th NODL=1 python tools/quick_perf.py 5000000 20 2>&1 | grep "iter 2"
RR_FAST_WARPS=$fw RR_FAST_THRESHOLD=$th NODL=1 python tools/quick_perf.py 5000000 200 2>&1 | grep "iter 2"
RR_FAST_WARPS=$fw RR_FAST_THRESHOLD=$th NODL=1 python tools/quick_perf.py 5000000 200 cfg4 2>&1 | grep "iter 2"
done; done
