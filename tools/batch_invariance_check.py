import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/self-similarity-grouping_b200")
import torch, ssg_b200
from ssg_b200 import synth
model = synth.build_model(2, 0)
imgs, _ = synth.synth_images(1300, 1234, torch.device("cuda", 0))
a = ssg_b200.embed_images(model, imgs, 2, False, 512, 0).clone()
b = ssg_b200.embed_images(model, imgs, 2, False, 96, 0).clone()
c = ssg_b200.embed_images(model, imgs[700:1300].contiguous(), 2, False, 512, 0).clone()
torch.cuda.synchronize()
d1 = (a - b).abs().amax(dim=(0, 2)); d2 = (a[:, 700:1300] - c).abs().amax(dim=(0, 2))
print(os.environ.get("TAG", "default"), "batch512 vs batch96: max", float(d1.max()), "rows differing", int((d1 > 0).sum()),
      "| offset slice: max", float(d2.max()), "rows differing", int((d2 > 0).sum()))
bad = torch.nonzero(d1 > 0).flatten()[:12].tolist(); print("first differing rows", bad)
