"""Drop-in installation into the reference package.

``install()`` rebinds the class names that the reference factory resolves from its own module globals
(deeplio/models/nets/__init__.py:6-10,95-102,141-144,173-176,206-209) to the B200-backed modules, so an
unmodified ``train.py`` / ``test.py`` / ``config.yaml`` build and run this path.  INTEGRATION.md shows the
two-line change a maintainer of the reference would make instead.
"""
import importlib

_NAMES = ("DeepLIO", "ImuFeatFC", "ImufeatRNN0", "LidarPointSegFeat", "LidarFlowNetFeat", "LidarSimpleFeat1",
          "LidarResNetFeat", "OdomFeatFC", "OdomFeatRNN", "DeepLIOFusionCat", "DeepLIOFusionSoft")


def install(target="deeplio.models.nets"):
    """Returns the list of rebound names.  ``target`` must already be importable."""
    from . import nets as ours
    ref = importlib.import_module(target)
    done = []
    for name in _NAMES:
        if hasattr(ref, name):
            setattr(ref, name, getattr(ours, name))
            done.append(name)
    return done
