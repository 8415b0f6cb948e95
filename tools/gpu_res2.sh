for shape in 4,512,8 8,1024,8 4,128,4 16,1024,8 8,256,8; do
echo "== shape $shape"
EB_PROBE_SHAPE=$shape python tools/res_probe.py 2>&1 | grep -v "niter="
done | tee gpurun_out/r02_res_probe_small.txt
