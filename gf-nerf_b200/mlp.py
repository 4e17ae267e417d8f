"""`MLPNetwork` with the constructor and parameter naming of the reference (gfnerf/mlp.py:3-57): a stack of
`nn.Linear` (+bias) with ReLU between layers, configured by a tiny-cuda-nn style dict.  The Linears live in
`self.layers = nn.ModuleList()` with contiguous indices exactly like the reference (gfnerf/mlp.py:35-43), so state
dicts are interchangeable with the reference's (`layers.<i>.weight / .bias`; tests/test_field_tables.py loads one
produced by the reference's own class).

On its own it is a parameter container: GF-NeRF never runs one of these stacks alone -- the density net and the
colour head are evaluated together by the fused tensor-core kernels (csrc/mlp.cu) through
`field.GFNeRFField.forward`, which reads the parameters of both stacks as one blob.  Calling `forward` on a single
stack therefore raises: there is deliberately no cuBLAS / PyTorch fallback on this path.
"""
import torch
from torch import nn


class MLPNetwork(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        self.n_input_dims, self.n_output_dims = int(n_input_dims), int(n_output_dims)
        self.n_neurons = int(network_config["n_neurons"])
        self.n_hidden_layers = int(network_config["n_hidden_layers"])
        self.activation = network_config.get("activation", "ReLU")
        self.output_activation = network_config.get("output_activation", "None")
        if self.activation != "ReLU":
            raise ValueError("gfnerf_b200 MLPNetwork: only ReLU hidden activations are built")
        self.layers = nn.ModuleList()
        d = self.n_input_dims
        for _ in range(self.n_hidden_layers):
            self.layers.append(nn.Linear(d, self.n_neurons))
            d = self.n_neurons
        self.layers.append(nn.Linear(d, self.n_output_dims))

    def linears(self):
        return list(self.layers)

    def flat_params(self) -> torch.Tensor:
        """weights then bias of every layer, nn.Linear layout -- the blob order of include/gfnerf_b200.h"""
        return torch.cat([torch.cat([l.weight.reshape(-1), l.bias.reshape(-1)]) for l in self.linears()])

    def forward(self, x):
        raise RuntimeError("gfnerf_b200.mlp.MLPNetwork is evaluated by the fused field kernel "
                           "(gfnerf_b200.field.GFNeRFField.forward); there is no stand-alone / PyTorch path")
