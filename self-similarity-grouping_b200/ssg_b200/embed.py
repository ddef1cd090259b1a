"""Embedding host API: ResNet-50 feature extraction with flip augmentation (reid/evaluators.py:18-60,
reid/feature_extraction/cnn.py:10-23, reid/models/resnet.py:86-134) on the CUDA trunk of libssg_b200."""
import ctypes
from collections import OrderedDict

import numpy as np

from . import _lib

_plans = {}
# the normaliser of the reference's loaders (selftraining.py:36-37)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def layer_table():
    """[(index, cin, cout, k, stride, conv_key, bn_key)] in the library's canonical order."""
    lib = _lib.load()
    out = []
    for i in range(lib.ssg_embed_num_layers()):
        cin, cout, k, s = (ctypes.c_int() for _ in range(4))
        ck, bk = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        _lib.check(lib.ssg_embed_layer_info(i, ctypes.byref(cin), ctypes.byref(cout), ctypes.byref(k),
                                            ctypes.byref(s), ck, bk, 64))
        out.append((i, cin.value, cout.value, k.value, s.value, ck.value.decode(), bk.value.decode()))
    return out


def unwrap(model):
    """DataParallel / DistributedDataParallel wrappers hand in ``.module`` (selftraining.py:135)."""
    while hasattr(model, "module") and not hasattr(model, "base"):
        model = model.module
    return model


def _check_num_split(num_split):
    """The pooled tail keeps up to 4 stripes + the global bank in registers (csrc/conv.cu pooled_tail_kernel)."""
    if not 1 <= int(num_split) <= 4:
        raise ValueError("ssg_b200: num_split=%r out of range (1..4)" % (num_split,))


class EmbedPlan(object):
    """Device workspace + folded weights of one ResNet-50 trunk for batches of up to ``batch_max`` images."""

    def __init__(self, batch_max=256, device=None, height=256, width=128):
        dev = _lib.require_cuda(device)
        self.device = dev
        self.batch_max = int(batch_max)
        self._h = ctypes.c_void_p()
        self._weights_token = None
        self._staging = {}           # host-image staging buffers / events of embed_images, per (image shape, dtype, batch)
        _lib.check(_lib.load().ssg_embed_plan_create(ctypes.byref(self._h), dev.index, self.batch_max, height, width))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                _lib.load().ssg_embed_plan_destroy(h)
            except Exception:
                pass

    @property
    def nbytes(self):
        return int(_lib.load().ssg_embed_plan_bytes(self._h))

    # ---- weights
    def load_state_dict(self, sd, prefix=""):
        """sd: torchvision-ResNet-50-style state_dict (keys ``conv1.weight``, ``layer1.0.bn1.running_var`` ...)."""
        import torch
        lib = _lib.load()
        for i, cin, cout, k, s, ck, bk in layer_table():
            def g(name):
                t = sd[prefix + name]
                return t.detach().to(self.device, dtype=torch.float32).contiguous()
            w = g(ck + ".weight")
            if tuple(w.shape) != (cout, cin, k, k):
                raise ValueError("layer %s: weight shape %s, expected %s" % (ck, tuple(w.shape), (cout, cin, k, k)))
            gamma, beta, mean, var = g(bk + ".weight"), g(bk + ".bias"), g(bk + ".running_mean"), g(bk + ".running_var")
            _lib.check(lib.ssg_embed_load_layer(self._h, i, w.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                                mean.data_ptr(), var.data_ptr(), 1e-5, _lib.stream_ptr(self.device)))
        torch.cuda.current_stream(self.device).synchronize()     # the staging tensors above die here

    def load_model(self, model):
        """Ingest ``model.base`` of a reference-style ResNet (re-ingests only when the parameters changed)."""
        m = unwrap(model)
        base = m.base
        token = tuple((p.data_ptr(), p._version) for p in base.parameters()) + \
            tuple((b.data_ptr(), b._version) for b in base.buffers())
        if token != self._weights_token:
            for mod in base.modules():
                eps = getattr(mod, "eps", None)
                if eps is not None and mod.__class__.__name__.startswith("BatchNorm") and abs(eps - 1e-5) > 1e-12:
                    raise ValueError("BatchNorm eps %g is not the torchvision default" % eps)
            self.load_state_dict(base.state_dict())
            self._weights_token = token
        return getattr(m, "num_split", 1)

    # ---- forward
    def forward(self, images, num_split=1, for_eval=False, flip=True, out=None, row0=0, mean=IMAGENET_MEAN,
                std=IMAGENET_STD):
        """images: float32 CUDA tensor [n,3,256,128], already normalised (what the reference's loader yields), or raw
        pixels as a uint8 CUDA tensor [n,256,128,3] that are normalised on the device with ``mean`` / ``std``
        (n <= batch_max).
        list mode  -> out [banks, rows, 2048] (bank b of image i at out[b, row0+i]);
        eval mode  -> out [rows, banks*2048]."""
        import torch
        assert images.is_cuda and images.dim() == 4
        if not _lib.same_device(images, self.device):
            raise ValueError("ssg_b200: images must live on the plan's device (%s), got %s" % (self.device, images.device))
        _check_num_split(num_split)
        images = images.contiguous()
        n = images.shape[0]
        banks = num_split + 1 if num_split > 1 else 1
        if out is None:
            out = torch.empty((n, banks * 2048) if for_eval else (banks, n, 2048), dtype=torch.float32,
                              device=images.device)
        bank_stride = 0 if for_eval else out.stride(0)
        if images.dtype == torch.uint8:
            if tuple(images.shape[1:]) != (256, 128, 3):
                raise ValueError("uint8 images must be HWC [n,256,128,3], got %s" % (tuple(images.shape),))
            c3 = ctypes.c_float * 3
            _lib.check(_lib.load().ssg_embed_forward_u8(self._h, images.data_ptr(), c3(*mean), c3(*std), n,
                                                        int(num_split), int(bool(for_eval)), int(bool(flip)),
                                                        out.data_ptr(), bank_stride, int(row0), _lib.stream_ptr(self.device)))
            return out
        if images.dtype != torch.float32 or tuple(images.shape[1:]) != (3, 256, 128):
            raise ValueError("images must be float32 [n,3,256,128] or uint8 [n,256,128,3], got %s %s"
                             % (images.dtype, tuple(images.shape)))
        _lib.check(_lib.load().ssg_embed_forward(self._h, images.data_ptr(), n, int(num_split), int(bool(for_eval)),
                                                 int(bool(flip)), out.data_ptr(), bank_stride, int(row0),
                                                 _lib.stream_ptr(self.device)))
        return out


    def forward_raw(self, images, num_split=1):
        """One forward without flip and without normalisation: [banks, n, 2048] pooled banks (cnn.py:16)."""
        import torch
        _check_num_split(num_split)
        images = images.contiguous()
        n = images.shape[0]
        banks = num_split + 1 if num_split > 1 else 1
        out = torch.empty((banks, n, 2048), dtype=torch.float32, device=images.device)
        _lib.check(_lib.load().ssg_embed_forward(self._h, images.data_ptr(), n, int(num_split), 2, 0,
                                                 out.data_ptr(), out.stride(0), 0, _lib.stream_ptr(self.device)))
        return out


def get_plan(batch_max=256, device=None):
    dev = _lib.require_cuda(device)
    plan = _plans.get(dev.index)
    if plan is None or plan.batch_max < batch_max:
        _plans.pop(dev.index, None)
        plan = EmbedPlan(max(batch_max, 256), dev.index)
        _plans[dev.index] = plan
    return plan


def embed_images(model, images, num_split=None, for_eval=False, batch=256, device=None, out_device=True,
                 mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None):
    """Embed a whole image tensor (normalised float32 [N,3,256,128] or raw uint8 [N,256,128,3]; host (ideally
    pinned) or device) in batches with copy/compute overlap.  Returns a CUDA tensor: [banks, N, 2048] (list mode)
    or [N, banks*2048] (eval mode); `out` (optional) receives the features instead of a fresh tensor."""
    import torch
    dev = _lib.require_cuda(device)
    plan = get_plan(batch, dev.index)
    ns = plan.load_model(model)
    num_split = ns if num_split is None else num_split
    banks = num_split + 1 if num_split > 1 else 1
    N = images.shape[0]
    if out is None:
        out = torch.empty((N, banks * 2048) if for_eval else (banks, N, 2048), dtype=torch.float32, device=dev)
    else:
        # caller-owned destination, e.g. this rank's slot of a multi-GPU gather buffer (rows may be strided per bank)
        want = (N, banks * 2048) if for_eval else (banks, N, 2048)
        if tuple(out.shape) != want or out.dtype != torch.float32 or not _lib.same_device(out, dev) or \
                out.stride(-1) != 1 or (not for_eval and out.stride(1) != 2048):
            raise ValueError("ssg_b200: `out` must be a float32 CUDA tensor of shape %s with contiguous rows" % (want,))
    if images.is_cuda:
        for r0 in range(0, N, batch):
            plan.forward(images[r0:r0 + batch], num_split, for_eval, True, out, r0, mean, std)
        return out
    # host images: double-buffered staging on a copy stream.  The staging buffers, their events and the stream live on
    # the plan and persist across calls: a slot is re-filled only after the forward that last read it has finished
    # (`freed`), INCLUDING a forward issued by the previous call.  (Round 1 allocated fresh buffers and a fresh stream
    # per call; the caching allocator handed the second call the first call's buffers while its last forwards were still
    # reading them, and the new copy stream overwrote them -- the last batches of the first set were embedded from
    # partly overwritten images.  Found in round 2 by comparing the host-image path with the device-image path.)
    compute = torch.cuda.current_stream(dev)
    key = (tuple(images.shape[1:]), images.dtype, int(batch))
    st = plan._staging.get(key)
    if st is None:
        st = {"stream": torch.cuda.Stream(device=dev),
              "bufs": [torch.empty((batch,) + tuple(images.shape[1:]), dtype=images.dtype, device=dev) for _ in range(2)],
              "ready": [torch.cuda.Event(), torch.cuda.Event()], "freed": [torch.cuda.Event(), torch.cuda.Event()],
              "used": [False, False]}
        st["stream"].wait_stream(compute)            # the allocations above are ordered on the compute stream
        plan._staging[key] = st
    copy_stream, bufs, ready, freed, used = st["stream"], st["bufs"], st["ready"], st["freed"], st["used"]
    starts = list(range(0, N, batch))
    for it, r0 in enumerate(starts):
        slot = it & 1
        n = min(batch, N - r0)
        with torch.cuda.stream(copy_stream):
            if used[slot]:
                copy_stream.wait_event(freed[slot])  # the forward that last read this slot (this call or an earlier one)
            bufs[slot][:n].copy_(images[r0:r0 + n], non_blocking=True)
            ready[slot].record(copy_stream)
        compute.wait_event(ready[slot])
        plan.forward(bufs[slot][:n], num_split, for_eval, True, out, r0, mean, std)
        freed[slot].record(compute)
        used[slot] = True
    return out


def extract_features(model, data_loader, print_freq=20, for_eval=True, metric=None):
    """Drop-in for reid/evaluators.py:18 extract_features (same signature, same return types):
    (OrderedDict fname -> CPU tensor | list of CPU tensors, OrderedDict fname -> pid)."""
    import time
    import torch
    model.eval()
    dev = _lib.require_cuda()
    m = unwrap(model)
    num_split = getattr(m, "num_split", 1)
    list_mode = (not for_eval) and num_split > 1
    banks = num_split + 1 if num_split > 1 else 1
    features, labels = OrderedDict(), OrderedDict()
    chunks, names, pids_all = [], [], []
    end = time.time()
    bt_sum, dt_sum, cnt = 0.0, 0.0, 0
    plan = None
    for i, (imgs, fnames, pids, cams) in enumerate(data_loader):
        dt = time.time() - end
        imgs = torch.as_tensor(imgs)
        if plan is None or plan.batch_max < imgs.shape[0]:
            plan = get_plan(max(256, imgs.shape[0]), dev.index)
            plan.load_model(model)
        # raw uint8 HWC batches (row f5: a loader that stops after Resize) are normalised on the device
        x = imgs.to(dev, non_blocking=True) if imgs.dtype == torch.uint8 else \
            imgs.to(dev, dtype=torch.float32, non_blocking=True)
        o = plan.forward(x, num_split, for_eval=not list_mode, flip=True)
        # list mode: [banks, n, 2048] (each bank normalised alone) -> [n, banks, 2048]; else [n, banks*2048]
        chunks.append(o.permute(1, 0, 2).reshape(o.shape[1], -1) if list_mode else o)
        names.extend(list(fnames))
        pids_all.extend(list(pids))
        bt = time.time() - end
        end = time.time()
        bt_sum += bt; dt_sum += dt; cnt += 1
        if (i + 1) % print_freq == 0:
            print('Extract Features: [{}/{}]\t'
                  'Time {:.3f} ({:.3f})\t'
                  'Data {:.3f} ({:.3f})\t'
                  .format(i + 1, len(data_loader), bt, bt_sum / cnt, dt, dt_sum / cnt))
    if not chunks:
        return features, labels
    flat = torch.cat(chunks, 0).cpu()                     # one device -> host copy for the whole set
    if list_mode:
        flat = flat.view(-1, banks, 2048)
        for k, (fname, pid) in enumerate(zip(names, pids_all)):
            features[fname] = [flat[k, b] for b in range(banks)]
            labels[fname] = pid
    else:
        for k, (fname, pid) in enumerate(zip(names, pids_all)):
            features[fname] = flat[k]
            labels[fname] = pid
    return features, labels
