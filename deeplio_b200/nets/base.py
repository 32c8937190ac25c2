"""Module protocol shared by all sub-nets, and parameter-container helpers.

``BaseNet`` mirrors the reference protocol (base_net.py:6-32): ``pretrained``, ``output_shape``,
``get_output_shape()``, ``name`` (lower-cased class name, used for checkpoint file names at
trainer.py:165-170), ``device`` and ``get_modules()``.

The encoders keep their weights in ordinary ``nn.Conv2d`` / ``nn.BatchNorm2d`` / ``nn.Linear`` objects
created under the reference's attribute names, so ``state_dict()`` keys, shapes, default initialisation,
``.to()``, ``requires_grad`` freezing and optimizer parameter order match the reference.  Those objects
are parameter containers only: their ``forward`` is never called -- compute goes through the C ABI.
"""
from torch import nn


class BaseNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.pretrained = False
        self.output_shape = None

    def get_output_shape(self):
        return self.output_shape

    @property
    def name(self):
        return self.__class__.__name__.lower()

    @property
    def device(self):
        devices = ({p.device for p in self.parameters()} | {b.device for b in self.buffers()})
        if len(devices) != 1:
            raise RuntimeError("Cannot determine device: {} different devices found".format(len(devices)))
        return next(iter(devices))

    def get_modules(self):
        return [self]


class ParamTree(nn.Module):
    """A container whose children are addressed by dotted names ('layer1.0.conv1')."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container: compute runs through the deeplio_b200 C ABI")

    def put(self, dotted, module):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            child = node._modules.get(part)
            if child is None:
                child = ParamTree()
                node.add_module(part, child)
            node = child
        node.add_module(parts[-1], module)
        return module


def conv_params(cin, cout, k, bias):
    kh, kw = (k, k) if isinstance(k, int) else k
    return nn.Conv2d(cin, cout, (kh, kw), padding=((kh - 1) // 2, (kw - 1) // 2), bias=bias)


def require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError("deeplio_b200: %s is on %s; the B200 path has no CPU fallback -- move the model and its "
                           "inputs to a CUDA device" % (what, t.device))
