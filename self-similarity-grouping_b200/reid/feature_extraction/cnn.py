"""Drop-in for reid/feature_extraction/cnn.py:10-23 extract_cnn_feature(model, inputs, for_eval, modules=None):
one no-grad forward of the trunk (no flip), outputs moved to the CPU (a list of per-bank tensors when the model
has num_split > 1 and for_eval is False, else one tensor)."""
import torch


def extract_cnn_feature(model, inputs, for_eval, modules=None):
    from ssg_b200 import _lib
    from ssg_b200.embed import get_plan, unwrap
    if modules is not None:
        raise NotImplementedError("forward-hook mode (cnn.py:25-35) is unused by the drivers and not provided")
    model.eval()
    dev = _lib.require_cuda()
    inputs = torch.as_tensor(inputs)
    m = unwrap(model)
    num_split = getattr(m, "num_split", 1)
    plan = get_plan(max(256, inputs.shape[0]), dev.index)
    plan.load_model(model)
    x = inputs.to(dev, dtype=torch.float32, non_blocking=True)
    list_mode = (not for_eval) and num_split > 1
    # un-normalised pooled banks: the normalisation belongs to extract_features (evaluators.py:32-35,42-43)
    out = plan.forward_raw(x, num_split)
    if list_mode:
        return [out[b].cpu() for b in range(out.shape[0])]
    return out.permute(1, 0, 2).reshape(out.shape[1], -1).cpu()
