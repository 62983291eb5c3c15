"""Import alias: the product package lives in ``kan-tts_b200/`` (a directory name
Python cannot import directly); ``import kantts_b200`` resolves to it."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "kan-tts_b200")
__path__[:] = [_real]
__file__ = _os.path.join(_real, "__init__.py")
exec(compile(open(__file__).read(), __file__, "exec"))
