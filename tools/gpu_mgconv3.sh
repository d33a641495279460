python tools/gpu_mgconv2.py 2>&1 | grep -v "coarse=300"
python tools/gpu_mgconv.py 2>&1 | grep -v resid
python tools/gpu_diag4.py 2>&1
