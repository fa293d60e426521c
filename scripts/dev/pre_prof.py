import sys, numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api
ctx = api.Context(0); st = torch.cuda.current_stream().cuda_stream
w, h, n = 640, 480, 216
rng = np.random.default_rng(1)
d_src = torch.from_numpy(rng.integers(0, 256, (8, h, w, 3), dtype=np.uint8)).cuda().repeat((n // 8, 1, 1, 1)).contiguous()
xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
fx = xs + 3.0 * np.sin(ys / 97.0) + 1.37; fy = ys + 2.0 * np.cos(xs / 131.0) - 0.61
m1 = np.stack([np.floor(fx), np.floor(fy)], -1).astype(np.int16)
m2 = ((np.floor((fy - np.floor(fy)) * 32).astype(np.uint16) << 5) | np.floor((fx - np.floor(fx)) * 32).astype(np.uint16)).astype(np.uint16)
d_m1, d_m2 = torch.from_numpy(m1).cuda(), torch.from_numpy(m2.view(np.int16)).cuda()
d_dst = torch.empty((n, h, w), dtype=torch.uint8, device="cuda")
for lighting in (1, 0):
    p = _abi.PreprocessParams(w, h, 3 * w, w, 3 * w * h, w * h, _abi.PRE_OPENCV4, lighting)
    for _ in range(2): ctx.preprocess_device(d_src.data_ptr(), n, d_m1.data_ptr(), d_m2.data_ptr(), p, d_dst.data_ptr(), st)
torch.cuda.synchronize()
