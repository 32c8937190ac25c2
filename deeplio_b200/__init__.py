"""deeplio_b200 -- B200-native (sm_100a) training hot path of DeepLIO behind the reference's model factory.

``deeplio_b200.nets.get_model(input_shape, cfg, device)`` is the drop-in for
``deeplio.models.nets.get_model`` (reference nets/__init__.py:16); ``deeplio_b200.install()`` rebinds the
reference's class registry so that its train.py / test.py run on this path unchanged.  All compute goes
through the C ABI in include/deeplio_b200.h (libdeeplio_b200.so); there is no CPU fallback.
"""
__version__ = "0.1.0"


def install():
    from .dropin import install as _install
    return _install()
