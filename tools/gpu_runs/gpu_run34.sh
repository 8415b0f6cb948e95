mkdir -p gpurun_out
for T in 32 64 128; do
echo "== T=$T"; timeout 120 tools/_build/microbench $T 4096 8 2>&1 | grep -E "eb_pt_swap|swap marks|swap    CTA|sharded swap, world=1, no" 
done | tee gpurun_out/micro_T.txt
