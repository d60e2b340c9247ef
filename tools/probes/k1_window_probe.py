"""Is the K1 area stage bound by the DRAM access pattern of the crop window?  The device-resident path reads 3312 of the
5760 bytes of every 1080p row (the columns the centre crop keeps); the host-upload path hands K1 the same window
COMPACTED (rows back to back).  Same kernel, same bytes, different layout: this prints the area-stage time of both."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from b200clip import capi
from b200clip import open_clip as oc

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
model, _, _ = oc.create_model_and_transforms("ViT-B-32", device="cuda:0", max_images=n, max_texts=1, seed=0)
h = model.handle
g = torch.Generator(device="cuda").manual_seed(0)
frames = torch.randint(0, 256, (n, 1080, 1920, 3), device="cuda", dtype=torch.uint8, generator=g)
host = torch.empty(n, 1080, 1920, 3, dtype=torch.uint8, pin_memory=True)
host.copy_(frames)
out = {}
for name, fn in (("device_resident_strided_window", lambda: model.encode_frames_u8(frames, capi.RESIZE_REFERENCE)),
                 ("host_upload_compacted_window", lambda: model.encode_frames_u8_host(host, capi.RESIZE_REFERENCE, True))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    h.profile_read(reset=True)
    h.profile_enable(True)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    h.profile_enable(False)
    prof = h.profile_read(reset=True)
    out[name] = {k: round(v["ms"] / 3, 4) for k, v in prof.items() if k.startswith("pre") and v["launches"]}
    out[name]["launches_pre_area"] = prof["pre_area"]["launches"] // 3
print(json.dumps({"frames": n, "k1_ms_per_pass": out}))
