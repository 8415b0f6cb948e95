for shape in "16 4096 8" "8 8192 8" "2 32768 8" "32 16384 20" "4 16384 20"; do
  echo "== $shape"
  timeout 120 tools/_build/microbench_prod $shape 2>&1 | grep -E "eb_stretch_step \(both|eb_pt_swap|eb_gauss|iteration"
done
