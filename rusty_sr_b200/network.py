"""Static description of the reference graph `sr_net(factor, None)` (src/network.rs:16-109):
which ops exist and in which order their parameters sit in the flat vector (op-registration
order, network.rs:27-76).
Pure metadata — the arithmetic is in csrc/."""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

CHANNELS = 3  # network.rs:13
FACTOR = 3    # main.rs:31


@dataclass(frozen=True)
class ParamSpan:
    name: str      # alumina op name (network.rs)
    kind: str      # "conv" | "bias" | "belu"
    shape: tuple   # conv: (Cout, Kh, Kw, Cin); others: (C,)
    offset: int
    line: int      # network.rs line that registers the op

    @property
    def count(self) -> int:
        n = 1
        for s in self.shape:
            n *= s
        return n


def param_layout(factor: int = FACTOR) -> List[ParamSpan]:
    ce = CHANNELS * factor * factor  # network.rs:37
    spec = [
        ("conv0", "conv", (32, 5, 5, 3), 33), ("f_bias", "bias", (32,), 34), ("f_activ", "belu", (32,), 35),
        ("expand_bias", "bias", (ce,), 38),
        ("l1_bias", "bias", (32,), 50), ("l2_bias", "bias", (32,), 51), ("l3_bias", "bias", (32,), 52),
        ("l1_activ", "belu", (32,), 54), ("l2_activ", "belu", (32,), 55), ("l3_activ", "belu", (32,), 56),
        ("conv1", "conv", (32, 5, 5, 32), 60), ("conv2", "conv", (32, 5, 5, 32), 61),
        ("conv3", "conv", (32, 5, 5, 32), 62),
        ("conv5", "conv", (32, 3, 3, 32), 65), ("conv6", "conv", (32, 3, 3, 32), 66),
        ("conv7", "conv", (ce, 3, 3, 32), 67),
        ("conv8", "conv", (32, 3, 3, 32), 69), ("conv9", "conv", (ce, 3, 3, 32), 70),
        ("conv10", "conv", (ce, 3, 3, 32), 72),
    ]
    out, off = [], 0
    for name, kind, shape, line in spec:
        span = ParamSpan(name, kind, shape, off, line)
        out.append(span)
        off += span.count
    return out


def num_params(factor: int = FACTOR) -> int:
    """graph.num_params() (main.rs:162): 130459 for factor 3."""
    spans = param_layout(factor)
    return spans[-1].offset + spans[-1].count
