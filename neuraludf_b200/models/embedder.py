"""Positional encoding, mirror of the reference's `models/embedder.py` (get_embedder :39-51).

Inside the networks the encoding is fused into the CUDA kernels (pe_forward_kernel / ew_pe_kernel); this module
keeps the reference's `get_embedder(multires, input_dims)` -> (embed_fn, out_dim) surface for external callers.
The stand-alone `embed_fn` is a few torch ops on whatever device its input lives on (it is not on the hot path).
"""
import torch


class Embedder:
    """[x | sin(2^0 x) | cos(2^0 x) | ... | sin(2^(L-1) x) | cos(2^(L-1) x)] (embedder.py:11-36)."""

    def __init__(self, **kwargs):
        self.kwargs = kwargs
        d = kwargs["input_dims"]
        n = kwargs["num_freqs"]
        max_freq = kwargs["max_freq_log2"]
        if kwargs.get("log_sampling", True):
            self.freq_bands = 2.0 ** torch.linspace(0.0, max_freq, n)
        else:
            self.freq_bands = torch.linspace(2.0 ** 0.0, 2.0 ** max_freq, n)
        self.include_input = kwargs.get("include_input", True)
        self.periodic_fns = kwargs.get("periodic_fns", [torch.sin, torch.cos])
        self.out_dim = (d if self.include_input else 0) + d * n * len(self.periodic_fns)

    def embed(self, inputs):
        parts = [inputs] if self.include_input else []
        for f in self.freq_bands.tolist():
            for fn in self.periodic_fns:
                parts.append(fn(inputs * f))
        return torch.cat(parts, -1)


def get_embedder(multires, input_dims=3):
    eo = Embedder(include_input=True, input_dims=input_dims, max_freq_log2=multires - 1, num_freqs=multires,
                  log_sampling=True, periodic_fns=[torch.sin, torch.cos])
    return (lambda x, eo=eo: eo.embed(x)), eo.out_dim
