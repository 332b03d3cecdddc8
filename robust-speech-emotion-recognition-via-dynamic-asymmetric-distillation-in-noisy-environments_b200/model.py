"""Drop-in model surface of the reference's I/model.py (identical in the CASIA / EMODB copies), with the encoder
running on the hand-written sm_100a kernels:

    Emotion2VecEncoder(x[B,T,768], padding_mask[B,T]|None) -> [B,256]      I/model.py:6-41
    EmotionClassifier(emb[B,256]) -> logits[B,4]                            I/model.py:44-64
    SSRLModel(cfg): student_/teacher_ encoder+classifier, load_complete_pretrained_weights, _init_teacher_network,
                    update_teacher_ema, predict, get_embeddings, ema_momentum            I/model.py:67-265

state_dict keys are the reference's 8 tensors ({student,teacher}_encoder.pre_net.{weight,bias},
{student,teacher}_classifier.fc_layer.{weight,bias}), so its .ckpt / .pth files load unchanged.
The training hot loop does not go through these modules (see engine.DADStepEngine); they serve validate / predict /
inference and any caller that wants the reference's autograd-style API."""
import torch
import torch.nn as nn

from . import ops


class Emotion2VecEncoder(nn.Module):
    """Linear(768->256) + ReLU + padding-masked mean over frames, fused in one kernel (no [B,T,256] tensor)."""

    def __init__(self, input_dim=768, hidden_dim=256, pretrained_path=None):
        super().__init__()
        if (input_dim, hidden_dim) != (ops.D_IN, ops.D_HID):
            raise ValueError(f"kernels are specialised for {ops.D_IN}->{ops.D_HID} (cfg.INPUT_DIM/HIDDEN_DIM), got {input_dim}->{hidden_dim}")
        self.pre_net = nn.Linear(in_features=input_dim, out_features=hidden_dim)
        self.activate = nn.ReLU()  # kept for module-tree compatibility; the activation is fused into the kernel

    def forward(self, x, padding_mask=None):
        ops.register_torch_ops()
        squeeze = x.dim() == 2
        if squeeze:
            x = x.unsqueeze(0)
            padding_mask = None if padding_mask is None else padding_mask.unsqueeze(0)
        has_pad = padding_mask is not None
        pad = padding_mask if has_pad else torch.empty(0, dtype=torch.bool, device=x.device)
        emb, _bits, _lens = torch.ops.dad.encoder_pool(x.float(), pad, self.pre_net.weight, self.pre_net.bias, has_pad)
        return emb


class EmotionClassifier(nn.Module):
    """Dropout + Linear(256->4).  [B,256]x[256,4] is a few kFLOP: left to ATen here; the training step fuses it with
    softmax / CE / KL in csrc/k_head.cu."""

    def __init__(self, input_dim: int, num_classes: int, dropout_rate: float = 0.1):
        super().__init__()
        self.dropout = nn.Dropout(dropout_rate)
        self.fc_layer = nn.Linear(in_features=input_dim, out_features=num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.fc_layer(self.dropout(x))


class SSRLModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        input_dim = getattr(cfg, "INPUT_DIM", 768)
        hidden_dim = getattr(cfg, "HIDDEN_DIM", 256)
        num_classes = getattr(cfg, "NUM_CLASSES", 4)
        dropout_rate = getattr(cfg, "DROPOUT_RATE", 0.1)
        pretrained_path = getattr(cfg, "PRETRAINED_EMOTION2VEC_PATH", None)
        self.student_encoder = Emotion2VecEncoder(input_dim, hidden_dim)
        self.student_classifier = EmotionClassifier(hidden_dim, num_classes, dropout_rate)
        self.teacher_encoder = Emotion2VecEncoder(input_dim, hidden_dim)
        self.teacher_classifier = EmotionClassifier(hidden_dim, num_classes, 0.0)  # teacher never drops (I/model.py:118-122)
        if pretrained_path:
            self.load_complete_pretrained_weights(pretrained_path)
        self._init_teacher_network()
        self.ema_momentum = getattr(cfg, "EMA_MOMENTUM", 0.99)

    def load_complete_pretrained_weights(self, pretrained_path):
        """Stage-1 checkpoint = bare state_dict with pre_net.* / post_net.* (I/model.py:143-198): pre_net -> student
        encoder, post_net -> student classifier fc_layer.  Like the reference, failures are reported and swallowed
        (training continues from the current weights)."""
        try:
            ckpt = torch.load(pretrained_path, map_location="cpu", weights_only=True)
            enc = {k: v for k, v in ckpt.items() if k.startswith("pre_net")}
            cls = {k.replace("post_net", "fc_layer"): v for k, v in ckpt.items() if k.startswith("post_net")}
            if enc:
                self.student_encoder.load_state_dict(enc, strict=False)
            if cls:
                self.student_classifier.load_state_dict(cls, strict=False)
            n = len(enc) + len(cls)
            print(f"pretrained weights: loaded {n}/4 tensors from {pretrained_path}" + ("" if n == 4 else " (incomplete: rest stays randomly initialised)"))
        except Exception as e:  # noqa: BLE001 -- reference behaviour (I/model.py:194-198)
            print(f"pretrained weights: loading failed ({e}); continuing with random initialisation")

    def _student_teacher_pairs(self):
        yield from zip(self.teacher_encoder.parameters(), self.student_encoder.parameters())
        yield from zip(self.teacher_classifier.parameters(), self.student_classifier.parameters())

    def _init_teacher_network(self):
        for t, s in self._student_teacher_pairs():
            t.data.copy_(s.data)
            t.requires_grad = False

    @torch.no_grad()
    def update_teacher_ema(self):
        """t = t*m + s*(1-m) with the reference's three roundings (I/model.py:211-223), one kernel per tensor."""
        for t, s in self._student_teacher_pairs():
            if t.is_cuda:
                ops.ema_(t.data.view(-1), s.data.contiguous().view(-1), self.ema_momentum)
            else:
                raise ops.DadError("update_teacher_ema: model is on the CPU; dad_b200 has no CPU path (move it to cuda)")

    def _fused_ok(self, x):
        return x.is_cuda and x.dim() == 3 and x.shape[1] >= 32 and x.shape[0] > 0

    def predict(self, x: torch.Tensor, padding_mask=None, use_teacher=False):
        """logits [B,4] in eval mode (I/model.py:225-245): encoder + classifier in one fused pass on the tcgen05 path (dad_predict)."""
        self.eval()  # the reference flips the WHOLE model to eval here (I/model.py:237); train_epoch flips it back
        with torch.no_grad():
            enc, cls = (self.teacher_encoder, self.teacher_classifier) if use_teacher else (self.student_encoder, self.student_classifier)
            if self._fused_ok(x):
                return ops.predict(x, padding_mask, enc.pre_net.weight, enc.pre_net.bias, cls.fc_layer.weight, cls.fc_layer.bias)["logits_s"]
            return cls(enc(x, padding_mask))  # 2-D input / utterances shorter than one 32-frame group: CUDA-core encoder kernel

    def predict_both(self, x: torch.Tensor, padding_mask=None):
        """(student logits, teacher logits) from ONE read of x (validate's disagreement rate, I/train.py:522-564, runs both
        networks over the same batches): the two weight sets share the converted A operand of one tcgen05 pass."""
        self.eval()
        with torch.no_grad():
            if self._fused_ok(x):
                se, sc, te, tc = self.student_encoder, self.student_classifier, self.teacher_encoder, self.teacher_classifier
                o = ops.predict(x, padding_mask, se.pre_net.weight, se.pre_net.bias, sc.fc_layer.weight, sc.fc_layer.bias,
                                te.pre_net.weight, te.pre_net.bias, tc.fc_layer.weight, tc.fc_layer.bias)
                return o["logits_s"], o["logits_t"]
            return (self.student_classifier(self.student_encoder(x, padding_mask)), self.teacher_classifier(self.teacher_encoder(x, padding_mask)))

    def get_embeddings(self, x: torch.Tensor, padding_mask=None, use_teacher=False):
        self.eval()
        with torch.no_grad():
            return (self.teacher_encoder if use_teacher else self.student_encoder)(x, padding_mask)
