for shape in "16 4096 8" "16 8192 8" "16 16384 8"; do
for v in 0 1 2 3; do
  echo "== $shape EB_SWAP_SKIP=$v"
  env EB_SWAP_SKIP=$v timeout 120 tools/_build/microbench $shape 2>&1 | grep -E "^eb_pt_swap|swap marks"
done; done
