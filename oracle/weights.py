"""Deterministic synthetic weights (test infrastructure).

`synth_state_dict(template)` fills every tensor of a state_dict from a per-key seeded CPU generator, so
the reference (dev container), the oracle and the CUDA implementation (GPU box) all load bit-identical
fp32 weights without shipping 500 MB checkpoints: the values depend only on (key, shape, seed) and on
torch's CPU Philox/MT stream, which is fixed for a given torch version (the same image runs everywhere).
"""
import hashlib

import torch


def _gen(key, seed):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:8], "little") % (2 ** 63))
    return g


def synth_tensor(key, shape, dtype=torch.float32, seed=0):
    g = _gen(key, seed)
    shape = tuple(shape)
    if key.endswith("num_batches_tracked"):
        return torch.zeros(shape, dtype=torch.long)
    if key.endswith("running_var"):
        return torch.rand(shape, generator=g) + 0.5
    if key.endswith("running_mean"):
        return torch.randn(shape, generator=g) * 0.1
    if key.endswith("grid"):  # KANLinear knot buffer: keep the template values
        return None
    if key.endswith("position_ids") or key.endswith("token_type_ids"):
        return None
    if key.endswith("layer_scale"):   # ConvNeXt CNBlock: the default 1e-6 would make every block the identity
        return torch.rand(shape, generator=g) * 0.3 + 0.1
    if len(shape) <= 1:
        if key.endswith("bn3.weight"):
            return torch.rand(shape, generator=g) * 0.2 + 0.1          # last BN of a bottleneck: damp the residual sum
        if key.endswith("weight") or key.endswith("spline_scaler"):
            return torch.rand(shape, generator=g) * 0.4 + 0.8          # norm scales around 1
        if key.endswith("w_img") or key.endswith("w_txt"):
            return torch.randn(shape, generator=g) * 0.5
        return torch.randn(shape, generator=g) * 0.02                   # biases
    if "embeddings" in key:
        return torch.randn(shape, generator=g) * 0.02
    if key.endswith("query"):
        return torch.randn(shape, generator=g)
    if key.endswith("w_gate") or key.endswith("w_noise"):
        return torch.randn(shape, generator=g) * 0.05
    fan_in = 1
    for s in shape[1:]:
        fan_in *= s
    gain = 2.0 if ("conv" in key or "downsample.0" in key) else 1.0    # ReLU convs keep activations O(1)
    if "cross_attention.query_conv" in key or "cross_attention.key_conv" in key:
        gain = 0.1   # ConNexT attention has no 1/sqrt(d): keep the 768-term logits O(1) so the softmax is not one-hot
    return torch.randn(shape, generator=g) * (gain / fan_in) ** 0.5


def synth_state_dict(template, seed=0):
    """template: a state_dict (only keys / shapes / dtypes are used).  Aliased keys of the reference's
    ImageEncoder (`stem.*`, `layerN.*` vs `model.*`) resolve to the same values."""
    out = {}
    for key, ref in template.items():
        canon = _canonical(key)
        t = synth_tensor(canon, ref.shape, ref.dtype, seed)
        out[key] = ref.detach().clone().cpu() if t is None else t.to(ref.dtype)
    return out


def _canonical(key):
    """image_encoder.stem.0.* == image_encoder.model.conv1.*, stem.1 == model.bn1, layerN == model.layerN."""
    if "image_encoder." in key and ".model." not in key:
        head, tail = key.split("image_encoder.", 1)
        if tail.startswith("stem.0."):
            return head + "image_encoder.model.conv1." + tail[len("stem.0."):]
        if tail.startswith("stem.1."):
            return head + "image_encoder.model.bn1." + tail[len("stem.1."):]
        if tail.startswith("layer"):
            return head + "image_encoder.model." + tail
    return key


def synthetic_batch(B, S, num_classes, seed=123, image_hw=224, unit_range=False):
    """Inputs of SURVEY.md section 8d: randn (or rand) images, ids with CLS=101, tail-padded mask, labels."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    images = torch.rand(B, 3, image_hw, image_hw, generator=g) if unit_range else torch.randn(B, 3, image_hw, image_hw, generator=g)
    ids = torch.randint(0, 30522, (B, S), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(min(8, S), S + 1, (B,), generator=g)
    mask = (torch.arange(S)[None, :] < lens[:, None]).long()
    labels = torch.randint(0, num_classes, (B,), generator=g)
    return images, ids, mask, labels
