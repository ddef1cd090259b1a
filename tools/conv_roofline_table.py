#!/usr/bin/env python
"""Per-launch roofline of the convolution trunk: joins the ncu per-launch table of one 512-image batch
(profiles/r02_final_conv_traffic.json: time, DRAM bytes) with the algorithmic flops of each launch and prints, per
launch, the tensor-bound and HBM-bound times against MEASURED_PEAKS.json and which of the two binds.
usage: python tools/conv_roofline_table.py > profiles/r02_conv_roofline_per_launch.md"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
prof = json.load(open(os.path.join(ROOT, "profiles", "r02_final_conv_traffic.json")))
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
TF, HBM = peaks["bf16_tflops_sustained"] * 1e12, peaks["hbm_gbs"] * 1e9
P = 2 * prof["batch_images"]                     # image-passes per batch (flip augmentation)

layers = []                                      # (name, M pixels per pass, [(K, N), ...] GEMMs fused in the launch)
layers.append(("stem 7x7/2 3->64 (+max-pool)", 8192, [(147, 64)]))
px = 2048
layers.append(("layer1.0 conv1 64->64", px, [(64, 64)]))
layers.append(("layer1.0 conv2 3x3 64->64", px, [(576, 64)]))
layers.append(("layer1.0 conv3+downsample (K 64|64) -> 256, + layer1.1 conv1 256->64", px, [(128, 256), (256, 64)]))
layers.append(("layer1.1 conv2 3x3", px, [(576, 64)]))
layers.append(("layer1.1 conv3 64->256 +res, + layer1.2 conv1", px, [(64, 256), (256, 64)]))
layers.append(("layer1.2 conv2 3x3", px, [(576, 64)]))
layers.append(("layer1.2 conv3 +res, + layer2.0 conv1 256->128", px, [(64, 256), (256, 128)]))
for li, (mid, cin_prev, px_in) in enumerate([(128, 256, 2048), (256, 512, 512), (512, 1024, 128)], start=2):
    px = px_in // 4
    nblocks = {2: 4, 3: 6, 4: 3}[li]
    if li > 2:
        layers.append(("layer%d.0 conv1 %d->%d" % (li, cin_prev, mid), px_in, [(cin_prev, mid)]))
    layers.append(("layer%d.0 conv2 3x3/2 %d->%d" % (li, mid, mid), px, [(9 * mid, mid)]))
    layers.append(("layer%d.0 conv3+downsample (K %d|%d) -> %d" % (li, mid, cin_prev, 4 * mid), px, [(mid + cin_prev, 4 * mid)]))
    for b in range(1, nblocks):
        layers.append(("layer%d.%d conv1 %d->%d" % (li, b, 4 * mid, mid), px, [(4 * mid, mid)]))
        layers.append(("layer%d.%d conv2 3x3 %d->%d" % (li, b, mid, mid), px, [(9 * mid, mid)]))
        layers.append(("layer%d.%d conv3 %d->%d +res" % (li, b, mid, 4 * mid), px, [(mid, 4 * mid)]))
assert len(layers) == len(prof["per_launch"]), (len(layers), len(prof["per_launch"]))

print("# Per-launch roofline of the convolution trunk (one 512-image batch = %d image-passes, round-2 final kernels)\n" % P)
print("Source: `profiles/r02_final_conv_traffic.json` (ncu per-launch time and DRAM bytes; cold-cache, serialised) joined with the")
print("algorithmic flops of each launch.  Peaks: MEASURED_PEAKS.json — %.1f TFLOP/s sustained bf16, %.1f GB/s HBM copy.\n"
      % (TF / 1e12, HBM / 1e9))
print("| # | launch | us | GFLOP | TFLOP/s | of tensor peak | DRAM MB | TB/s | of HBM peak | bound | roofline us | us / roofline |")
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|")
tot_t = tot_roof = tot_tensor = tot_hbm = 0.0
n_hbm = 0
for i, ((name, m, gemms), row) in enumerate(zip(layers, prof["per_launch"])):
    fl = sum(2.0 * P * m * k * n for k, n in gemms)
    by = row["dram_read"] + row["dram_write"]
    t = row["us"] * 1e-6
    t_tensor, t_hbm = fl / TF, by / HBM
    roof = max(t_tensor, t_hbm)
    bound = "tensor" if t_tensor >= t_hbm else "HBM"
    n_hbm += bound == "HBM"
    tot_t += t; tot_roof += roof; tot_tensor += t_tensor; tot_hbm += t_hbm
    print("| %d | %s | %.1f | %.1f | %.0f | %.2f | %.0f | %.2f | %.2f | %s | %.1f | %.2f |"
          % (i, name, row["us"], fl / 1e9, fl / t / 1e12, fl / t / TF, by / 1e6, by / t / 1e12, by / t / HBM, bound, roof * 1e6, t / roof))
print("\n**Totals per batch**: measured %.0f us; tensor-only bound %.0f us (flops / sustained bf16 peak: the `roofline.frac` of "
      "bench.py = %.2f under ncu); HBM-only bound %.0f us; per-launch roofline (max of the two, no cross-layer fusion beyond what "
      "the launches already fuse) %.0f us -> the trunk runs at **%.2f of its per-launch roofline**; %d of %d launches are HBM-bound "
      "on that roofline." % (tot_t * 1e6, tot_tensor * 1e6, tot_tensor / tot_t, tot_hbm * 1e6, tot_roof * 1e6, tot_roof / tot_t,
                             n_hbm, len(layers)))
