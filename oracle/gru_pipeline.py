"""TEST INFRASTRUCTURE ONLY - CPU restatement of what happens to the filter output on its way into the GRU, and the
consumer network itself, for BASELINE config 5 (end-to-end parity of the batched-KF arm vs the reference-KF arm).

    feature rows   data_collection/data_conversion_Kalman_to_Training.py:245-254
    min-max        gru/gru_train.py:56-63,108-111
    windows        gru/gru_train.py:180-192 (sequence_length = 10, gru_train.py:32)
    consumer       gru/gru_model.py:7-49  RNN(188, 128, 4, 24): 4-layer GRU, last step, Linear, sigmoid
The trained weights and the ViT latents are not shipped: weights come from torch.manual_seed(1),
the 128-d latent is seeded U(0,1), identical in both arms.
"""
from __future__ import annotations

import numpy as np
import torch


def feature_rows(x, imu_acc, f, p_world, dp, imu):
    """All arguments [T, C] for one trajectory -> [T, 60] in the driver's column order."""
    return np.concatenate([x, imu_acc, f, p_world, dp, imu], axis=1)


def min_max_normalise(rows):
    lo, hi = rows.min(axis=0), rows.max(axis=0)
    return (rows - lo) / (hi - lo), lo, hi


def windows(norm_rows, latent, seq_len=10):
    """norm_rows [R, 60] float64, latent [R, 128] -> float32 [R - seq_len + 1, seq_len, 188] (torch.tensor(..., float32))."""
    full = np.concatenate([norm_rows, latent.astype(np.float64)], axis=1)
    out = np.stack([full[i:i + seq_len] for i in range(full.shape[0] - seq_len + 1)])
    return out.astype(np.float32)


class ConsumerRNN(torch.nn.Module):
    """Same architecture and parameter names as the reference's gru_model.RNN (state_dicts are interchangeable)."""

    def __init__(self, input_size=188, hidden_size=128, num_layers=4, num_classes=24):
        super().__init__()
        self.num_layers, self.hidden_size = num_layers, hidden_size
        self.gru = torch.nn.GRU(input_size, hidden_size, num_layers, batch_first=True)
        self.fc = torch.nn.Linear(hidden_size, num_classes)

    def forward(self, x):
        h0 = torch.zeros(self.num_layers, x.size(0), self.hidden_size, device=x.device, dtype=x.dtype)
        out, _ = self.gru(x, h0)
        return torch.sigmoid(self.fc(out[:, -1, :]))


def seeded_consumer():
    torch.manual_seed(1)
    return ConsumerRNN().eval()


def seeded_latent(n_rows, n_latent=128, seed=7):
    return np.random.default_rng(seed).uniform(0.0, 1.0, (n_rows, n_latent)).astype(np.float32)


def per_state_rmse(pred_norm, truth_rows, seq_len=10):
    """GRU output [:, :12] is a min-max normalised state; de-normalise with the label's range and compare with the label of the
    window's last row (gru_test.py:208-217,341-369, without its 10-sample KF offset)."""
    lo, hi = truth_rows.min(axis=0), truth_rows.max(axis=0)
    span = np.where(hi > lo, hi - lo, 1.0)
    pred = pred_norm[:, :12] * span + lo
    lab = truth_rows[seq_len - 1:]
    return np.sqrt(((pred - lab) ** 2).mean(axis=0))
