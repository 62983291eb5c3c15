"""Import the UNMODIFIED reference (/root/reference) in the build container
(TEST INFRASTRUCTURE; used only by tests/golden/make_golden.py and by tests
that are skipped when /root/reference is absent, e.g. on the GPU box).

The reference imports several packages that are not installed in this image;
none of them is on the hot path, except pytorch_wavelets.DWT1DForward and
librosa.filters.mel, which are replaced by the restatements in oracle/dwt.py
and oracle/melbasis.py (so those two stay "parity unpinned", see
oracle/__init__.py).  Recipe from SURVEY.md section 8(c).
"""
import os
import sys
import types

REF_ROOT = os.environ.get("KANTTS_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "kantts"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_shims():
    if getattr(install_shims, "_done", False):
        return
    from . import dwt, melbasis

    if "pytorch_wavelets" not in sys.modules:
        _stub("pytorch_wavelets", DWT1DForward=dwt.DWT1DForward)
    if "librosa" not in sys.modules:
        def _mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw):
            return melbasis.mel_filterbank(sr, n_fft, n_mels, fmin, fmax)
        filters = _stub("librosa.filters", mel=_mel)
        _stub("librosa", filters=filters, __version__="0.9.2")
        _stub("librosa.util")
        _stub("librosa.effects")
    if "ttsfrd" not in sys.modules:
        _stub("ttsfrd")

    class _SummaryWriter:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def add_figure(self, *a, **k):
            pass

        def add_audio(self, *a, **k):
            pass

        def close(self):
            pass

    if "tensorboardX" not in sys.modules:
        _stub("tensorboardX", SummaryWriter=_SummaryWriter)
    for name in ("soundfile", "unidecode", "inflect", "pysptk", "sox", "bitstring"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                _stub(name, unidecode=lambda s: s, engine=lambda: None)
    if "matplotlib" not in sys.modules:
        try:
            import matplotlib  # noqa: F401
        except Exception:
            plt = _stub("matplotlib.pyplot")
            _stub("matplotlib", pyplot=plt, use=lambda *a, **k: None)
    import scipy.signal
    if not hasattr(scipy.signal, "kaiser"):
        scipy.signal.kaiser = scipy.signal.windows.kaiser
    install_shims._done = True


def import_reference():
    """-> the ``kantts`` package of the unmodified reference."""
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    install_shims()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import warnings
    warnings.filterwarnings("ignore")
    import kantts  # noqa: F401
    import kantts.models  # noqa: F401
    import kantts.train.loss  # noqa: F401
    return kantts
