"""GPU, marker `gpu_next` (NOT part of `-m gpu`): opt-in variants written after the round's GPU budget was spent.
Run `python -m pytest tests -m gpu_next -q` on a B200 before making any of them a default; each must reproduce the
default path bit for bit."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _embed_in_subprocess(tmp_path, name, env, n_img=21, batch=32):
    script = (
        "import sys, numpy as np, torch\n"
        "sys.path[:0] = %r\n"
        "import ssg_b200\n"
        "from oracle import resnet_oracle as R\n"
        "plan = ssg_b200.EmbedPlan(%d); plan.load_model(R.build_model(2, 0))\n"
        "out = plan.forward(R.synth_images(%d, 11).cuda(), 2); torch.cuda.synchronize()\n"
        "np.save(sys.argv[1], out.cpu().numpy())\n"
        % ([os.path.join(ROOT, "self-similarity-grouping_b200"), ROOT], batch, n_img))
    out_file = str(tmp_path / (name + ".npy"))
    subprocess.run([sys.executable, "-c", script, out_file], check=True, env=dict(os.environ, **env), timeout=600)
    return np.load(out_file)


@pytest.mark.gpu_next
@pytest.mark.parametrize("chunk", [8, 10, 32])
def test_l2_chunked_layers_are_bit_identical(tmp_path, chunk):
    """SSG_L2_CHUNK: layers 1-2 over chunks of image-passes that stay in L2 (embed.cu) -- same kernels on the same
    per-image tiles, so the features must equal the unchunked forward bit for bit (21 images = 42 passes: chunk 8 and
    10 leave a ragged last chunk, 32 a short one)."""
    want = _embed_in_subprocess(tmp_path, "plain", {"SSG_L2_CHUNK": "0"})
    got = _embed_in_subprocess(tmp_path, "chunk%d" % chunk, {"SSG_L2_CHUNK": str(chunk)})
    assert np.isfinite(want).all()
    assert np.array_equal(got, want)
