for v in "EB_STRETCH_HT=256" "EB_STRETCH_HT=128" "EB_STRETCH_HT=64" "EB_STRETCH_SPLIT=1"; do
  echo "== $v"
  env $v EB_STRETCH_VERBOSE=1 timeout 120 tools/_build/microbench 16 4096 8 2>&1 | grep -E "max active|eb_stretch_step \(both|stretch CTA" | sort | uniq | head -5
done
