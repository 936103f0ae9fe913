"""TEST INFRASTRUCTURE (checker only; nothing under cutter-vad_b200/ imports this).

Silero VAD v5, 16 kHz branch, re-assembled from PyTorch's OWN modules (torch.nn.functional.conv1d, torch.nn.LSTMCell)
with the tensors the reference's silero_vad_v5.onnx carries under their original torch names
(`If_0_then_branch__Inline_0__<name>`: `stft.forward_basis_buffer`, `encoder.N.reparam_conv.{weight,bias}`,
`decoder.rnn.{weight,bias}_{ih,hh}`, `decoder.decoder.2.{weight,bias}`).

Why it exists: onnxruntime is absent from this image, so the numpy op-by-op interpretation of the ONNX file
(oracle/onnx_interp.py) is what every parity test hangs on.  This module is an independent second opinion on that
interpretation from third-party kernels: the graph was exported from exactly such a PyTorch module, its LSTM node
re-orders torch's (i, f, g, o) gate rows into ONNX's (i, o, f, c) and back (SURVEY.md section 8a stage 9), and if the
interpreter mis-read any of that, `torch.nn.LSTMCell` fed the file's raw `weight_ih / weight_hh` would disagree with it.
The wiring (pad 64 right reflect, k=256 s=128 STFT conv, magnitude, four conv + ReLU, LSTMCell, ReLU -> 1x1 conv ->
sigmoid) follows the reference's call site `silero_model.py:403-537` fed 512 samples and SURVEY.md section 8a "S5".
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn.functional as F


class TorchV5:
    def __init__(self, named: Dict[str, np.ndarray], dtype=torch.float32):
        t = lambda k: torch.from_numpy(np.ascontiguousarray(named[k])).to(dtype)   # noqa: E731
        self.dtype = dtype
        self.basis = t("stft.forward_basis_buffer")                      # [258, 1, 256]
        self.enc = [(t(f"encoder.{i}.reparam_conv.weight"), t(f"encoder.{i}.reparam_conv.bias"), s)
                    for i, s in enumerate((1, 2, 2, 1))]
        self.cell = torch.nn.LSTMCell(128, 128).to(dtype)
        with torch.no_grad():
            self.cell.weight_ih.copy_(t("decoder.rnn.weight_ih"))
            self.cell.weight_hh.copy_(t("decoder.rnn.weight_hh"))
            self.cell.bias_ih.copy_(t("decoder.rnn.bias_ih"))
            self.cell.bias_hh.copy_(t("decoder.rnn.bias_hh"))
        self.dec_w = t("decoder.decoder.2.weight")                       # [1, 128, 1]
        self.dec_b = t("decoder.decoder.2.bias")

    @torch.no_grad()
    def frame(self, x: np.ndarray, h: np.ndarray, c: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
        """x [B, 512], h / c [B, 128] -> (prob [B], h', c')."""
        xt = torch.from_numpy(np.ascontiguousarray(x)).to(self.dtype)[:, None, :]
        xt = F.pad(xt, (0, 64), mode="reflect")                          # /stft/padding: right side only
        st = F.conv1d(xt, self.basis, stride=128)                        # [B, 258, 3]
        e = torch.sqrt(st[:, :129] ** 2 + st[:, 129:] ** 2)
        for w, b, s in self.enc:
            e = F.relu(F.conv1d(e, w, b, stride=s, padding=1))
        assert e.shape[2] == 1
        hn, cn = self.cell(e[:, :, 0], (torch.from_numpy(h).to(self.dtype), torch.from_numpy(c).to(self.dtype)))
        p = torch.sigmoid(F.conv1d(F.relu(hn)[:, :, None], self.dec_w, self.dec_b))[:, 0, 0]
        return p.numpy().astype(np.float64), hn.numpy(), cn.numpy()

    def run(self, audio: np.ndarray, n_frames: int, hop: int = 512, frame_len: int = 512, denoise: bool = False):
        """audio [B, L] -> probs [B, n_frames] with the reference's framing / gate / zero-pad around the model."""
        B = audio.shape[0]
        h = np.zeros((B, 128), np.float32)
        c = np.zeros((B, 128), np.float32)
        out = np.zeros((B, n_frames))
        for j in range(n_frames):
            f = audio[:, j * hop:j * hop + frame_len].astype(np.float32)
            if denoise:
                f = np.where(np.abs(f) > 0.01, f, 0.0).astype(np.float32)          # audio.py:117-118
            if frame_len < 512:
                f = np.pad(f, ((0, 0), (0, 512 - frame_len)))                      # silero_model.py:464-468
            out[:, j], h, c = self.frame(f[:, :512], h, c)
        return out
