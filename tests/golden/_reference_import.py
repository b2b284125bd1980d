"""Import shim used ONLY by tests/golden/make_reference_golden.py (run in the build container where /root/reference
exists).  The reference targets pydantic 1.10 and a dozen packages that are not installed here; this module aliases
`pydantic` to the bundled `pydantic.v1` for the reference's own modules and installs inert stub modules for the
missing third-party imports that the code under test never calls (open_clip, pycurl, nltk, ...).  Nothing here is
shipped or imported by marqo_b200/."""
import builtins
import importlib
import sys
import types

REFERENCE_SRC = "/root/reference/src"
STUB_ROOTS = ['open_clip', 'clip', 'sentence_transformers', 'semver', 'more_itertools', 'nltk', 'validators', 'pycurl',
              'ftfy', 'kazoo', 'readerwriterlock', 'onnxruntime', 'timm', 'magic', 'ffmpeg', 'cachetools', 'redis',
              'onnx', 'decord', 'pytorchvideo', 'torchaudio', 'pynvml', 'memory_profiler', 'cv2', 'multilingual_clip',
              'kornia', 'boto3', 'botocore', 'jinja2']


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        m = _Stub(self.__name__ + '.' + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return _Stub('call')

    def __mro_entries__(self, bases):
        return (object,)


def install():
    # real packages that need the real pydantic v2 must be imported BEFORE the alias
    import fastapi  # noqa: F401
    import starlette  # noqa: F401
    import transformers  # noqa: F401
    import torch  # noqa: F401
    import torchvision  # noqa: F401
    try:
        import httpx  # noqa: F401
    except ImportError:
        STUB_ROOTS.append('httpx')
    import pydantic.v1 as pv1
    sys.modules['pydantic'] = pv1
    for sub in ('error_wrappers', 'fields', 'main', 'validators', 'types', 'errors', 'typing', 'utils',
                'class_validators', 'generics', 'schema'):
        sys.modules['pydantic.' + sub] = importlib.import_module('pydantic.v1.' + sub)
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    real_import = builtins.__import__
    stubbed = []

    def fake_import(name, globals=None, locals=None, fromlist=(), level=0):
        try:
            return real_import(name, globals, locals, fromlist, level)
        except ImportError:
            root = name.split('.')[0]
            if level == 0 and root in STUB_ROOTS:
                parts = name.split('.')
                for i in range(1, len(parts) + 1):
                    n = '.'.join(parts[:i])
                    if n not in sys.modules:
                        sys.modules[n] = _Stub(n)
                stubbed.append(name)
                return sys.modules[name if fromlist else root]
            raise

    builtins.__import__ = fake_import
    return stubbed
