#!/usr/bin/env python
"""LRU model of the DRAM traffic of the embedding trunk's convolution launches, per batch (analysis aid, no GPU).

Why: the 1x1 convolutions of layers 1-2 are HBM bound and SSG_L2_CHUNK (embed.cu) runs those layers
over chunks of image-passes that re-use the head of the activation buffers so that they stay in the 126 MB L2.  This
script replays the trunk's launch sequence (embed.cu run_block: conv1 -> conv2 -> conv3 [+ K-concatenated downsample in
the first block of a layer]) at the granularity of 32 KB pieces of the NHWC bf16 activation tensors through a
write-back, write-allocate LRU cache and counts DRAM reads (misses) and writes (dirty evictions):
  * `cold`    : cache flushed before every launch -- what `ncu` measures (its default --cache-control all); compared
                with profiles/r01_final_conv_traffic.json to validate the byte accounting;
  * `default` : the launch sequence as it runs (whole batch per launch), warm cache;
  * `chunk=c` : layers 1-2 over chunks of c image-passes in chunk-local buffers, layers 3-4 over the whole batch.
Not modelled: weights (<= 2.4 MB per launch, L2 hits after the first tile), the stem, set associativity, the two L2
partitions of the two dies (bracketed by running the model at 126 and at 63 MB), TMA halo re-reads (L2 hits).

    python tools/l2_traffic_model.py [--batch 512] [--md profiles/r01_l2_chunk_model.md]
"""
import argparse
import json
import os
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = 32 * 1024                      # granule
BLOCKS = (3, 4, 6, 3)


class Cache(object):
    def __init__(self, cap_bytes):
        self.cap = cap_bytes // G
        self.d = OrderedDict()
        self.rd = self.wr = 0

    def read(self, key):
        if key in self.d:
            self.d.move_to_end(key)
        else:
            self.rd += G
            self._put(key, False)

    def write(self, key):
        if key in self.d:
            self.d[key] = True
            self.d.move_to_end(key)
        else:
            self._put(key, True)               # full-line TMA stores: no fill read

    def _put(self, key, dirty):
        self.d[key] = dirty
        if len(self.d) > self.cap:
            _, was_dirty = self.d.popitem(last=False)
            if was_dirty:
                self.wr += G

    def flush(self):
        for dirty in self.d.values():
            if dirty:
                self.wr += G
        self.d.clear()


def trunk_launches():
    """[(layer, block, name, reads [(buffer, bytes per image-pass, fraction of the tensor touched)], write (buffer, bytes))]
    buffers: 'x' / 'y' ping-pong, 't1', 't2'; sizes in bytes per image-pass (bf16 NHWC)."""
    out = []
    px, cin = 64 * 32, 64
    xb, yb = "x", "y"
    for L in range(4):
        mid, outc = 64 << L, 256 << L
        for b in range(BLOCKS[L]):
            stride = 2 if (b == 0 and L > 0) else 1
            opx = px // (stride * stride)
            out.append((L, b, "conv1", [(xb, px * cin * 2, 1.0)], ("t1", px * mid * 2)))
            out.append((L, b, "conv2", [("t1", px * mid * 2, 1.0)], ("t2", opx * mid * 2)))
            if b == 0:
                # [t2 | x] . [W3 | Wds]^T: the stride-2 view of x touches one pixel in four (whole 128 B+ pixel rows)
                out.append((L, b, "conv3+ds", [("t2", opx * mid * 2, 1.0), (xb, px * cin * 2, 1.0 / (stride * stride))],
                            (yb, opx * outc * 2)))
            else:
                out.append((L, b, "conv3", [("t2", opx * mid * 2, 1.0), (xb, opx * outc * 2, 1.0)], (yb, opx * outc * 2)))
            xb, yb = yb, xb
            px, cin = opx, outc
    return out


def pieces(nbytes, frac=1.0):
    n = max(1, int(round(nbytes * frac / G)))
    return n


def run_launch(cache, launch, passes, base, tag):
    """passes: image-passes of this launch; base[buffer] -> address offset (in passes) of pass 0; tag distinguishes
    buffers that live at different addresses (chunk-local vs full-batch)."""
    _, _, _, reads, (wbuf, wbytes) = launch
    for p in range(passes):
        for buf, nbytes, frac in reads:
            for g in range(pieces(nbytes, frac)):
                cache.read((tag.get(buf, buf), base.get(buf, 0) + p, g))
        for g in range(pieces(wbytes)):
            cache.write((tag.get(wbuf, wbuf), base.get(wbuf, 0) + p, g))


def simulate(nb, cap_mb, chunk=0, cold=False):
    cache = Cache(int(cap_mb * 1024 * 1024))
    launches = trunk_launches()
    per = []                                   # (layer, name, read, write) per launch of the default sequence

    def account(launch, passes, base, tag):
        r0, w0 = cache.rd, cache.wr
        run_launch(cache, launch, passes, base, tag)
        if cold:
            cache.flush()
        per.append((launch[0], launch[1], launch[2], cache.rd - r0, cache.wr - w0))

    if not chunk:
        for l in launches:
            account(l, nb, {}, {})
    else:
        # the pooled stem map is produced for the whole batch first; its tail is still in L2 when the chunk loop starts,
        # the head has been written back: model it as not resident
        l12 = [l for l in launches if l[0] < 2]
        l34 = [l for l in launches if l[0] >= 2]
        first_in = l12[0][3][0][0]             # 'x': the chunk's slice of the full-batch pooled map
        last_out = l12[-1][4][0]               # buffer name the last layer-2 block writes
        for q0 in range(0, nb, chunk):
            c = min(chunk, nb - q0)
            for i, l in enumerate(l12):
                tag, base = {}, {}
                if i < 3 and True:
                    pass
                # block (0, 0) reads the full-batch pooled map at pass offset q0 (conv1 and the fused downsample)
                if l[0] == 0 and l[1] == 0:
                    tag[first_in] = "pooled_all"
                    base["x"] = q0
                if i == len(l12) - 1:
                    tag[last_out] = "l2_all"
                    base[last_out] = q0
                account(l, c, base, tag)
        for i, l in enumerate(l34):
            tag = {}
            if l[0] == 2 and l[1] == 0:
                tag[l[3][0][0]] = "l2_all"     # layer 3 starts from the full-batch layer-2 output
            account(l, nb, {}, tag)
    cache.flush()                              # the features leave through the pooled tail; count what is still dirty
    by_layer = {}
    for L, b, name, r, w in per:
        a = by_layer.setdefault(L, [0, 0])
        a[0] += r
        a[1] += w
    return cache.rd, cache.wr, by_layer, per


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=512, help="images per embedding batch (two passes each)")
    ap.add_argument("--md", default=None)
    args = ap.parse_args()
    nb = 2 * args.batch
    lines = []
    rd, wr, by_layer, per = simulate(nb, 126, cold=True)
    meas = None
    path = os.path.join(ROOT, "profiles", "r01_final_conv_traffic.json")
    if os.path.isfile(path) and args.batch == 512:
        with open(path) as f:
            tj = json.load(f)
        # launch 0 of the capture is the stem; the remaining 48 are the trunk's launches in order
        meas = tj["per_launch"][1:]
    lines.append("# LRU model of the convolution launches' DRAM traffic, batch of %d images (%d image-passes)\n" % (args.batch, nb))
    lines.append("Made by `tools/l2_traffic_model.py` (no GPU).  Granule 32 KB, write-back / write-allocate LRU; weights, stem, "
                 "associativity and the split of the L2 over the two dies are not modelled (the last is bracketed by the 63 MB rows).\n")
    lines.append("## 1. Validation of the byte accounting: cold cache per launch vs the ncu capture\n")
    lines.append("`ncu` flushes the caches before every launch, so its per-launch DRAM bytes are the model's `cold` mode "
                 "(`profiles/r01_final_conv_traffic.json`, launches 1-48; launch 0 is the stem).\n")
    lines.append("| layer | model read + write, GB | ncu read + write, GB |")
    lines.append("|---|---:|---:|")
    if meas and len(meas) == len(per):
        ml = {}
        for (L, b, name, r, w), m in zip(per, meas):
            a = ml.setdefault(L, 0.0)
            ml[L] = a + m["dram_read"] + m["dram_write"]
        for L in range(4):
            lines.append("| %d | %.2f | %.2f |" % (L + 1, sum(by_layer[L]) / 1e9, ml[L] / 1e9))
        lines.append("| all | %.2f | %.2f |" % ((rd + wr) / 1e9, sum(ml.values()) / 1e9))
    else:
        for L in range(4):
            lines.append("| %d | %.2f | n/a |" % (L + 1, sum(by_layer[L]) / 1e9))
    lines.append("")
    lines.append("## 2. Warm cache: the default launch sequence against L2-resident chunks of layers 1-2\n")
    lines.append("| schedule | L2 modelled, MB | DRAM read, GB | DRAM write, GB | total, GB | layers 1-2, GB | layers 3-4, GB |")
    lines.append("|---|---:|---:|---:|---:|---:|---:|")
    for cap in (126, 63):
        for chunk in (0, 8, 16, 32, 48, 64):
            rd, wr, by_layer, _ = simulate(nb, cap, chunk=chunk)
            l12 = sum(sum(by_layer[L]) for L in (0, 1))
            l34 = sum(sum(by_layer[L]) for L in (2, 3))
            lines.append("| %s | %d | %.2f | %.2f | %.2f | %.2f | %.2f |" % ("default" if not chunk else "SSG_L2_CHUNK=%d" % chunk, cap,
                                                                         rd / 1e9, wr / 1e9, (rd + wr) / 1e9, l12 / 1e9, l34 / 1e9))
    text = "\n".join(lines) + "\n"
    print(text)
    if args.md:
        with open(args.md, "w") as f:
            f.write(text)


if __name__ == "__main__":
    main()
