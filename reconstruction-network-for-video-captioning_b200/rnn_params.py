"""Parameter holder with nn.LSTM / nn.GRU's state_dict layout (weight_ih_l{k}, weight_hh_l{k}, bias_*).

The reference stores its recurrent weights inside nn.LSTM / nn.GRU (models/decoder.py:36-40).  We keep the
same names, shapes, gate order and initialiser (uniform(-1/sqrt(H), 1/sqrt(H)) in parameter order, so the
same torch seed gives the same draws) but never run a cuDNN / ATen RNN: the weights are read by our kernels.
"""
import math

import torch
import torch.nn as nn


class RNNParams(nn.Module):
    def __init__(self, model_name: str, input_size: int, hidden_size: int, num_layers: int = 1, dropout: float = 0.0):
        super().__init__()
        self.mode = "LSTM" if model_name == "LSTM" else "GRU"      # models/decoder.py:32-35: anything else is GRU
        self.input_size, self.hidden_size, self.num_layers, self.dropout = input_size, hidden_size, num_layers, dropout
        gates = 4 if self.mode == "LSTM" else 3
        for layer in range(num_layers):
            in_l = input_size if layer == 0 else hidden_size
            self.register_parameter(f"weight_ih_l{layer}", nn.Parameter(torch.empty(gates * hidden_size, in_l)))
            self.register_parameter(f"weight_hh_l{layer}", nn.Parameter(torch.empty(gates * hidden_size, hidden_size)))
            self.register_parameter(f"bias_ih_l{layer}", nn.Parameter(torch.empty(gates * hidden_size)))
            self.register_parameter(f"bias_hh_l{layer}", nn.Parameter(torch.empty(gates * hidden_size)))
        self.reset_parameters()

    def reset_parameters(self):
        stdv = 1.0 / math.sqrt(self.hidden_size) if self.hidden_size > 0 else 0
        for w in self.parameters():
            nn.init.uniform_(w, -stdv, stdv)

    def layer(self, k: int = 0):
        return (getattr(self, f"weight_ih_l{k}"), getattr(self, f"weight_hh_l{k}"),
                getattr(self, f"bias_ih_l{k}"), getattr(self, f"bias_hh_l{k}"))

    def forward(self, *a, **k):
        raise RuntimeError("RNNParams only holds weights; the recurrence runs in recnet_b200's CUDA kernels")

    def extra_repr(self):
        return f"{self.mode}, {self.input_size} -> {self.hidden_size}, layers={self.num_layers}"
