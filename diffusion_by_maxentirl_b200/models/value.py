"""Drop-in for the reference's models/value.py::TimeIndependentValue (:3-14)."""
import torch.nn as nn


class TimeIndependentValue(nn.Module):
    def __init__(self, net):
        super().__init__()
        self.net = net

    def forward(self, x, t, y=None, out=None):
        if out is not None:  # B200 extension: write the energies into a caller-provided buffer (packed gather)
            return self.net(x, y, out=out)
        return self.net(x, y) if y is not None else self.net(x)

    def load_pretrained(self, ckpt):
        """Reference value.py:14-15."""
        self.net.load_pretrained(ckpt)
