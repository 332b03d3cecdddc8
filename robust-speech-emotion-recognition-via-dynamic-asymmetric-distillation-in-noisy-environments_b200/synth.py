"""Deterministic synthetic DAD batches: class-structured 768-d "emotion2vec-like" frame
features, a noisy domain built with the white-noise SNR power formula of the reference's offline script
(pretrain-and-processed-*/scripts/add_noise_to_audio.py:14-43, without its peak renormalisation) applied in
feature space, and the per-step RNG inputs of the parity-mode step.

Batch dict layout = the reference collator's (I/dataload_clean.py:177-193, I/dataload_noisy.py:111-129).
Everything is generated on the CPU generator so tests, golden fixtures and the bench see identical bytes."""
import torch

D_IN, D_HID, N_CLS = 768, 256, 4


def class_means(seed=1234):
    g = torch.Generator().manual_seed(seed)
    return 0.5 * torch.randn(N_CLS, D_IN, generator=g)


def _lengths(B, T, g, variable):
    if not variable:
        return torch.full((B,), T, dtype=torch.long)
    lens = torch.randint(max(1, T // 2), T + 1, (B,), generator=g)
    lens[torch.randint(0, B, (1,), generator=g)] = T  # collator pads to the batch max
    return lens


def make_batches(B, T, seed, snr_db=10.0, variable_len=False, mu=None):
    """-> (clean_batch, noisy_batch) dicts.  padding_mask True = pad; padded frames are zero."""
    g = torch.Generator().manual_seed(seed)
    mu = class_means() if mu is None else mu
    out = []
    for domain in ("clean", "noisy"):
        y = torch.randint(0, N_CLS, (B,), generator=g)
        x = mu[y][:, None, :] + torch.randn(B, T, D_IN, generator=g)
        lens = _lengths(B, T, g, variable_len)
        pad = torch.arange(T)[None, :] >= lens[:, None]
        if domain == "noisy":
            # fp64 accumulation, rounded to fp32 once: the result does not depend on the host's SIMD reduction order,
            # so fixtures generated in the build container reproduce bit-for-bit on the GPU box's CPU
            valid = (~pad).unsqueeze(-1).double()
            xd = x.double()
            P = (xd * xd * valid).sum(dim=(1, 2)) / (lens.double() * D_IN)
            sigma = torch.sqrt(P / (10.0 ** (snr_db / 10.0))).float()
            x = x + sigma[:, None, None] * torch.randn(B, T, D_IN, generator=g)
        x = x.masked_fill(pad.unsqueeze(-1), 0.0).contiguous()
        out.append({"id": torch.arange(B), "net_input": {"feats": x, "padding_mask": pad}, "labels": y})
    return out[0], out[1]


def make_rng_inputs(B, T, seed, drop_p=0.1, mask_ratio=0.1):
    """Supplied draws for one step, in the reference's draw order."""
    g = torch.Generator().manual_seed(seed)
    L = int(T * mask_ratio)
    return {
        "drop_clean": (torch.rand(B, D_HID, generator=g) >= drop_p).float(),
        "n_weak": torch.randn(B, T, D_IN, generator=g),
        "n_strong": torch.randn(B, T, D_IN, generator=g),
        "feat_keep_u": torch.rand(D_IN, generator=g),
        "t_start": torch.randint(0, max(1, T - L + 1), (B,), generator=g),
        "drop_strong": (torch.rand(B, D_HID, generator=g) >= drop_p).float(),
    }


def init_state(seed=42):
    """Default nn.Linear init of the 4 student tensors (kaiming-uniform a=sqrt(5) == U(+-1/sqrt(fan_in)) for
    both weight and bias), teacher = copy (I/model.py:200-209).  Keys = SSRLModel.state_dict()."""
    g = torch.Generator().manual_seed(seed)

    def lin(o, i):
        k = 1.0 / (i ** 0.5)
        return (torch.rand(o, i, generator=g) * 2 - 1) * k, (torch.rand(o, generator=g) * 2 - 1) * k

    W, b = lin(D_HID, D_IN)
    Wc, bc = lin(N_CLS, D_HID)
    s = {"student_encoder.pre_net.weight": W, "student_encoder.pre_net.bias": b,
         "student_classifier.fc_layer.weight": Wc, "student_classifier.fc_layer.bias": bc}
    s.update({k.replace("student_", "teacher_"): v.clone() for k, v in list(s.items())})
    return s
