"""throughput of the pre-pass (usv_preprocess_device) with frames resident in HBM, against the HBM roofline.
Algorithmic bytes per frame: 3*W*H read (BGR) + W*H written (gray); the rectification maps (6 B/pixel) are shared by the
whole batch and counted once. CUDA events on the launching stream, 3 warm-up launches, batches larger than the 126 MB L2.
Also times OpenCV's CPU implementation of the same chain (cv2, all host threads) on a bounded sample of the batch."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, ".")
import torch
from unsynchronized_stereo_vision_proj325_b200 import _abi, api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]; SRC = "measured"
except Exception:
    HBM, SRC = 6650.0, "fallback"
ctx = api.Context(0)
st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(325)


def maps_for(w, h):
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    fx = xs + 3.0 * np.sin(ys / 97.0) + 1.37           # a smooth warp of a few pixels, like a rectification
    fy = ys + 2.0 * np.cos(xs / 131.0) - 0.61
    m1 = np.stack([np.floor(fx), np.floor(fy)], -1).astype(np.int16)
    m2 = ((np.floor((fy - np.floor(fy)) * 32).astype(np.uint16) << 5) | np.floor((fx - np.floor(fx)) * 32).astype(np.uint16)).astype(np.uint16)
    return m1, m2


def cpu_chain(frames, m1, m2, lighting):
    import cv2
    cv2.setNumThreads(0)
    t0 = time.perf_counter()
    for f in frames:
        r = cv2.remap(f, m1, m2, cv2.INTER_LINEAR, borderMode=cv2.BORDER_CONSTANT, borderValue=0)
        if lighting:
            hsv = cv2.cvtColor(r, cv2.COLOR_BGR2HSV)
            hsv[..., 2] = cv2.equalizeHist(np.ascontiguousarray(hsv[..., 2]))
            r = cv2.cvtColor(hsv, cv2.COLOR_HSV2BGR)
        cv2.cvtColor(r, cv2.COLOR_BGR2GRAY)
    return time.perf_counter() - t0


for (w, h, n) in ((640, 480, 1024), (1920, 1080, 128)):
    base = rng.integers(0, 256, (8, h, w, 3), dtype=np.uint8)
    d_src = torch.from_numpy(base).cuda().repeat((n // 8, 1, 1, 1)).contiguous()
    m1, m2 = maps_for(w, h)
    d_m1, d_m2 = torch.from_numpy(m1).cuda(), torch.from_numpy(m2.view(np.int16)).cuda()
    pitch = -(-w // 16) * 16
    d_dst = torch.empty((n, h, pitch), dtype=torch.uint8, device="cuda")
    for lighting in (1, 0):
        p = _abi.PreprocessParams(w, h, 3 * w, pitch, 3 * w * h, pitch * h, _abi.PRE_OPENCV4, lighting)
        run = lambda: ctx.preprocess_device(d_src.data_ptr(), n, d_m1.data_ptr(), d_m2.data_ptr(), p, d_dst.data_ptr(), st)
        l0 = ctx.launch_count
        for _ in range(3): run()
        per_call = (ctx.launch_count - l0) // 3
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        reps = 10
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        algo = n * w * h * 4 + w * h * 6
        k = min(n, 16 if w <= 640 else 4)
        host = base[:8][np.arange(k) % 8]
        t_cpu = cpu_chain(host, m1, m2, bool(lighting))
        print(json.dumps({"what": "pre-pass: remap + %sBGR2GRAY (P/Main.cpp:914-921)" % ("BGR2HSV + equalizeHist(V) + HSV2BGR + " if lighting else ""),
                          "frames": n, "frame": [w, h], "ms_per_batch": ms, "frames_per_s": n / ms * 1e3, "gpu_launches_per_batch": per_call,
                          "roofline": {"bound": "hbm", "achieved": algo / ms / 1e6, "peak": HBM, "unit": "GB/s", "frac": algo / ms / 1e6 / HBM,
                                       "peak_source": SRC, "algorithmic_bytes_per_frame": w * h * 4, "traffic": None},
                          "cpu_baseline": {"value": k / t_cpu, "unit": "frames/s", "kind": "cv2 %s (OpenCV CPU, all host threads)" % __import__("cv2").__version__,
                                           "cores": os.cpu_count(), "sample": "%d frames" % k}}), flush=True)
