"""``AcousticSystem`` (acoustic_system.py:3-51): defender -> transform -> classifier, with the int16-range renormalisation.
The reference class itself also works unchanged with the audiopure_b200 modules; this copy of the interface exists so the
package is self-contained (the GPU box has no reference checkout)."""
from __future__ import annotations

import torch

__all__ = ["AcousticSystem"]


class AcousticSystem(torch.nn.Module):
    def __init__(self, classifier: torch.nn.Module, transform, defender: torch.nn.Module = None, defense_type: str = "wave",
                 check_int16_range: bool = True):
        super().__init__()
        # the range test reads two reductions back to the host (as the reference does); switch it off to make forward()
        # free of host synchronisation, e.g. for CUDA-graph capture of small-batch query serving
        self.check_int16_range = check_int16_range
        self.classifier = classifier
        self.transform = transform
        self.defender = defender
        self.defense_type = defense_type
        if self.defense_type not in ["wave", "spec"]:
            raise NotImplementedError("argument defense_type should be 'wave' or 'spec'!")

    def forward(self, x, defend=True):
        # int16-range input -> [-1, 1)                                   (acoustic_system.py:29-30)
        if self.check_int16_range:
            lo, hi = torch.aminmax(x)
            if 0.9 * hi > 1 and 0.9 * lo < -1:
                x = x / (2 ** 15)
        use_defender = defend is True and self.defender is not None
        output = self.defender(x) if (use_defender and self.defense_type == "wave") else x
        if self.transform is not None:
            output = self.transform(output)
        if use_defender and self.defense_type == "spec":
            output = self.defender(output)
        return self.classifier(output)
