"""Seeded synthetic inputs for the parity tests, the golden generator and bench.py.

All generators take an explicit seed and run on the CPU generator so that the same call gives the
same tensors in the build container (where the goldens are frozen) and on the GPU box.
SURVEY.md section 8d describes the distributions.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch


def gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


# ---- CenterNet -----------------------------------------------------------------------------------

def centernet_model_config(in_h=512, in_w=512, downsamples=2):
    ratio = 2 ** downsamples
    return SimpleNamespace(in_h=in_h, in_w=in_w, downsamples=downsamples, downsample_ratio=ratio,
                           out_h=in_h // ratio, out_w=in_w // ratio)


def separated_logits(B, C, H, W, seed, lo=-6.0, hi=3.0):
    """Every frame is a random permutation of a strictly increasing ramp: all logits are distinct, and the
    top of the ramp (the only part a top-k can reach) is spaced 1e-4 apart, i.e. tens of fp32 ulps apart even
    after the sigmoid, so the ranked order is unambiguous whatever the last-ulp rounding of exp()."""
    g = gen(seed)
    n = C * H * W
    m = min(n // 2, 20000)
    top = hi - 1e-4 * torch.arange(m - 1, -1, -1, dtype=torch.float64)
    rest = torch.linspace(lo, float(top[0]) - 1e-4, n - m, dtype=torch.float64) if n > m else top[:0]
    ramp = torch.cat((rest, top)).to(torch.float32)
    assert ramp.numel() == n
    out = torch.empty((B, n), dtype=torch.float32)
    for b in range(B):
        out[b] = ramp[torch.randperm(n, generator=g)]
    return out.reshape(B, C, H, W)


def natural_logits(B, C, H, W, seed, n_peaks=(8, 32)):
    """N(-2.2, 1.5^2) background (the reference's heatmap bias init is -2.19) with planted Gaussian
    bumps of sigma in [1,4] at random classes."""
    g = gen(seed)
    x = torch.randn((B, C, H, W), generator=g) * 1.5 - 2.2
    yy, xx = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    for b in range(B):
        k = int(torch.randint(n_peaks[0], n_peaks[1] + 1, (1,), generator=g))
        for _ in range(k):
            c = int(torch.randint(0, C, (1,), generator=g))
            cy = float(torch.rand((1,), generator=g)) * (H - 1)
            cx = float(torch.rand((1,), generator=g)) * (W - 1)
            s = 1.0 + 3.0 * float(torch.rand((1,), generator=g))
            amp = 4.0 + 4.0 * float(torch.rand((1,), generator=g))
            x[b, c] += amp * torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return x


def head_views(B, H, W, seed, with_depth=True):
    """size / offset / depth exactly as Centernet.forward hands them out: NCHW tensors viewed NHWC
    (centernet.py:81-89), i.e. non-contiguous."""
    g = gen(seed)
    size = (torch.rand((B, 2, H, W), generator=g) * 0.3 + 0.02).permute(0, 2, 3, 1)
    offset = (torch.rand((B, 2, H, W), generator=g) * 4.0).permute(0, 2, 3, 1)
    depth = (torch.randn((B, 1, H, W), generator=g)).permute(0, 2, 3, 1) if with_depth else None
    return size, offset, depth


def pose_truth(B, n_obj, C, seed, n_kp_inst=0, Kp=0, p_valid=0.75):
    """PoseSample-shaped truth (datasets/load/pose_dataset.py:24-41); only the fields the encoders read."""
    g = gen(seed)
    t = SimpleNamespace()
    t.img = None
    t.valid = torch.rand((B, n_obj), generator=g) < p_valid
    t.label = torch.randint(0, C, (B, n_obj), generator=g)
    t.center = torch.rand((B, n_obj, 2), generator=g)
    t.size = torch.rand((B, n_obj, 2), generator=g) * 0.3
    if n_kp_inst:
        t.keypoint_valid = torch.rand((B, n_kp_inst), generator=g) < p_valid
        t.keypoint_label = torch.randint(0, Kp, (B, n_kp_inst), generator=g)
        t.keypoint_center = torch.rand((B, n_kp_inst, 2), generator=g)
        t.keypoint_object_index = torch.randint(0, n_obj, (B, n_kp_inst), generator=g)
    return t


def truth_to(t, device):
    out = SimpleNamespace()
    for k, v in vars(t).items():
        setattr(out, k, v.to(device) if isinstance(v, torch.Tensor) else v)
    return out


# ---- YOLACT --------------------------------------------------------------------------------------

def yolact_config(in_h=550, in_w=550, scales=(24, 48, 96, 192, 384), ratios=(0.5, 1, 2), variances=(0.1, 0.2),
                  pos=0.4, neg=0.3, negative_example_ratio=3):
    return SimpleNamespace(in_h=in_h, in_w=in_w, anchor_scales=scales, anchor_aspect_ratios=ratios,
                           box_variances=variances, iou_pos_threshold=pos, iou_neg_threshold=neg,
                           negative_example_ratio=negative_example_ratio)


def fpn_sizes(in_h, in_w):
    """resnet18 strides 8/16/32 then two stride-2 3x3 convs with padding 1 (backbone.py:21-23,
    feature_pyramid.py:22-25,55-56): 550 -> 69, 35, 18, 9, 5."""
    def conv(n, k, s, p):
        return (n + 2 * p - k) // s + 1
    h, w = in_h, in_w
    h, w = conv(h, 7, 2, 3), conv(w, 7, 2, 3)      # stem
    h, w = conv(h, 3, 2, 1), conv(w, 3, 2, 1)      # maxpool
    sizes = []
    h, w = conv(h, 3, 2, 1), conv(w, 3, 2, 1)      # layer2 (stride 8)
    sizes.append((h, w))
    for _ in range(4):
        h, w = conv(h, 3, 2, 1), conv(w, 3, 2, 1)
        sizes.append((h, w))
    return sizes


def yolact_heads(B, N, C1, seed, anchor, n_clusters=12, per_cluster=12, separated=False):
    """Class logits + box encodings with planted confident priors in clusters (so NMS has work to do).

    separated=False: cls ~ N(0, 2^2), background-dominated, cluster members boosted — "natural" data whose
    confidences can tie to within an ulp.
    separated=True : every prior's max-foreground confidence is sigmoid(margin) for a distinct margin from a
    shuffled ramp (all other foreground logits are -30, i.e. below fp32 resolution of the softmax sum), with
    the largest margins given to the cluster members.  Ranked order is then unambiguous whatever the
    last-ulp rounding of exp() / the softmax summation order."""
    g = gen(seed)
    enc = torch.randn((B, N, 4), generator=g) * 0.5
    if not separated:
        cls = torch.randn((B, N, C1), generator=g) * 2.0
        cls[:, :, 0] += 4.0  # background-dominated like a trained head
    else:
        cls = torch.full((B, N, C1), -30.0)
        cls[:, :, 0] = 0.0
    for b in range(B):
        members_all = []
        for c in range(n_clusters):
            centre = int(torch.randint(0, N, (1,), generator=g))
            members = (centre + torch.randperm(min(N, 64), generator=g)[:per_cluster]) % N
            enc[b, members] *= 0.2
            members_all.append(members)
            if not separated:
                k = int(torch.randint(1, C1, (1,), generator=g))
                cls[b, members, k] += 8.0 + 4.0 * torch.rand((per_cluster,), generator=g)
                cls[b, members, 0] -= 4.0
        if separated:
            ramp = torch.linspace(-8.0, 4.0, N)
            members = torch.unique(torch.cat(members_all))
            others = torch.tensor(sorted(set(range(N)) - set(members.tolist())), dtype=torch.int64)
            order = torch.cat((others[torch.randperm(others.numel(), generator=g)],
                               members[torch.randperm(members.numel(), generator=g)]))
            margin = torch.empty(N)
            margin[order] = ramp  # members receive the top of the ramp
            k = torch.randint(1, C1, (N,), generator=g)
            cls[b, torch.arange(N), k] = margin
    return cls, enc


def mask_inputs(P, H, W, K, seed):
    g = gen(seed)
    proto = torch.nn.functional.leaky_relu(torch.randn((P, H, W), generator=g))
    coeff = torch.tanh(torch.randn((K, P), generator=g))
    box = torch.cat((torch.rand((K, 2), generator=g) * 0.8 + 0.1, torch.rand((K, 2), generator=g) * 0.5 + 0.05), dim=-1)
    return proto, coeff, box


def depth_image(hi, wi, seed, holes=0.2):
    """A mono16 depth image (millimetres) with `holes` of the pixels carrying no reading (0), as int32 on the host
    (torch has no CPU arithmetic on uint16); .to(torch.uint16) gives the device-side form."""
    g = gen(seed)
    d = torch.randint(300, 9000, (hi, wi), generator=g, dtype=torch.int32)
    return torch.where(torch.rand((hi, wi), generator=g) < holes, torch.zeros((), dtype=torch.int32), d)


def mask_inputs_exact(P, H, W, K, seed):
    """Prototypes / coefficients that are small integers (exact in bf16, sums exact in fp32) with half-integer logits:
    every implementation must select exactly the same pixels."""
    g = gen(seed)
    proto = torch.randint(-3, 4, (P, H, W), generator=g).float()
    proto[0] = 1.0
    coeff = torch.randint(-2, 3, (K, P), generator=g).float()
    coeff[:, 0] = 0.5
    box = torch.cat((torch.rand((K, 2), generator=g) * 0.8 + 0.1, torch.rand((K, 2), generator=g) * 0.7 + 0.1), dim=-1)
    return proto, coeff, box


def truth_boxes(B, M, seed):
    g = gen(seed)
    box = torch.cat((torch.rand((B, M, 2), generator=g) * 0.8 + 0.1, torch.rand((B, M, 2), generator=g) * 0.4 + 0.05),
                    dim=-1)
    valid = torch.rand((B, M), generator=g) < 0.75
    return box, valid
