python -m pytest tests/test_deflate_gpu.py tests/test_stream_gpu.py -m gpu -x -q 2>&1 | tail -2
python - <<'PY'
import sys; sys.path.insert(0,'.')
from zlib_b200 import load
import torch
L=load(); assert L.dll.zb200_init(0)==0
from zlib_b200 import binding as zb
n=1<<30
src=torch.from_numpy(L.synth(n,kind=1,seed=1)).cuda(); cap=L.compress_bound(n)+64; dst=torch.empty(cap,dtype=torch.uint8,device='cuda')
s=torch.cuda.current_stream()
for lv in (1,2,3,6):
    L.deflate(src.data_ptr(),n,dst.data_ptr(),cap,lv,zb.WRAP_ZLIB,s); torch.cuda.synchronize()
    L.profile(True); c=L.deflate(src.data_ptr(),n,dst.data_ptr(),cap,lv,zb.WRAP_ZLIB,s); r=L.profile_report(); L.profile(False)
    print(lv, n/c, {k:round(v[0],2) for k,v in r.items() if 'walk' in k})
PY
python tools/probe_codec.py 1024 2>&1 | head -1
