"""B200-native likelihood-tempered Sequential Monte Carlo (hot path only).

Drop-in for the sampler loop of
`maruchitatsuki/python-based-Sequential-Monte-Carlo-method-with-likelihood-tempering`
(per-particle log-likelihood -> tempering -> residual-systematic resampling -> MH mutation).
Python host code calls hand-written sm_100a CUDA kernels in `libsmcb200.so` through ctypes
(`include/smcb200.h`); torch only owns device tensors and provides `torch.distributed`.

The directory name contains hyphens, so import it with
`importlib.import_module("python-based-sequential-monte-carlo-method-with-likelihood-tempering_b200")`
or through the `smcb200` alias module at the repository root.
"""
from . import _lib
from .settings import Settings
from .prior import UniformBox, IndependentPrior
from .likelihood import MMProgress, MMRate, KineticRK, KineticDAE, UserKernelLikelihood, CallableLikelihood
from .artefacts import RunWriter

__all__ = ["Settings", "UniformBox", "IndependentPrior", "MMProgress", "MMRate", "KineticRK", "KineticDAE", "UserKernelLikelihood", "CallableLikelihood",
           "RunWriter", "Engine", "run", "build",
           "LocalComm", "TorchComm", "NcclComm", "migration_plan"]


def build(force=False):
    """Compile libsmcb200.so (nvcc, sm_100a) if it is missing or stale."""
    from . import _build
    return _build.build(force=force)


def build_user_library(src, out=None, force=False):
    """nvcc -> shared library for a user-written likelihood kernel (see include/smcb_user.cuh)."""
    from . import _build
    return _build.build_user_library(src, out, force)


def __getattr__(name):
    # engine imports torch; keep `import package` light for build-only use
    if name in ("Engine", "LocalComm", "TorchComm", "NcclComm", "migration_plan", "Result", "StageRecord"):
        from . import engine
        return getattr(engine, name)
    if name == "run":
        return run
    raise AttributeError(name)


def run(likelihood, prior, particles=None, settings=None, comm=None, stream=None, **kw):
    """One call for the whole sampler: a likelihood, a prior, prior particles (or None to draw them
    on the device) and settings in; `Result` (posterior particles, beta schedule, log-evidence) out."""
    from .engine import Engine
    eng = Engine(likelihood, prior, settings, comm=comm)
    try:
        if particles is None:
            eng.sample_prior()
        return eng.run(particles, stream=stream, **kw)
    finally:
        eng.close()
