python tools/gpu_mgconv2.py
echo "--- pad=2"; FDFD_MG_PAD=2 python tools/gpu_mgconv2.py 2>&1 | grep -E "levels=8 coarse=4|512x256|proto"
echo "--- pad=4"; FDFD_MG_PAD=4 python tools/gpu_mgconv2.py 2>&1 | grep -E "levels=8 coarse=4|512x256|proto"
