"""Import the reference's own modules (models, utils, data, train) for the oracle.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

Sources of the modules, in order: /root/reference (build container), else the byte-compiled copies under
oracle/_ref/ (oracle/build_ref.py; they travel to the GPU box).  `nltk` and `sacred` are not installed in this
image and are not on the hot path: minimal stand-ins are registered so `data.py` / `train.py` import (the Sacred
decorators become identities, so train.eval_link_prediction is a plain function).
"""
import importlib.machinery
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("BLP_REFERENCE", "/root/reference")
REF_PYC = os.path.join(HERE, "_ref")
_cache = {}


def install_stubs():
    if "nltk" not in sys.modules:
        nltk = types.ModuleType("nltk")
        nltk.download = lambda *a, **k: True
        nltk.word_tokenize = lambda s: s.split()
        corpus = types.ModuleType("nltk.corpus")
        corpus.stopwords = types.SimpleNamespace(words=lambda lang: [])
        nltk.corpus = corpus
        sys.modules.update({"nltk": nltk, "nltk.corpus": corpus})
    if "sacred" not in sys.modules:
        sacred = types.ModuleType("sacred")

        class Experiment:
            def __init__(self, *a, **k):
                self.observers = []
                self.logger = None

            def _ident(self, f):
                return f
            config = capture = command = automain = main = _ident

            def run_commandline(self, *a, **k):
                return None
        sacred.Experiment = Experiment
        run_mod = types.ModuleType("sacred.run")
        run_mod.Run = object
        obs = types.ModuleType("sacred.observers")
        obs.MongoObserver = object
        sys.modules.update({"sacred": sacred, "sacred.run": run_mod, "sacred.observers": obs})


def available():
    """'source', 'compiled' or None."""
    if os.path.isfile(os.path.join(REF_SRC, "models.py")):
        return "source"
    if os.path.isfile(os.path.join(REF_PYC, "models.bin")):
        return "compiled"
    return None


def load(names=("models", "utils")):
    """dict name -> module, the reference's own code; raises ImportError when neither form is present.
    The modules are registered in sys.modules under their reference names (train.py does `import models, utils`)."""
    kind = available()
    if kind is None:
        raise ImportError("reference modules not available: neither /root/reference nor oracle/_ref/*.bin "
                          "(run `python oracle/build_ref.py` where the reference is mounted)")
    install_stubs()
    out = {}
    order = [n for n in ("models", "utils", "data", "train") if n in names or
             (n in ("models", "utils", "data") and "train" in names) or
             (n == "models" and "utils" in names)]                      # utils.py does `import models`
    for name in order:
        if name in _cache:
            out[name] = _cache[name]
            continue
        if kind == "source":
            path = os.path.join(REF_SRC, name + ".py")
            spec = importlib.util.spec_from_file_location(name, path)
        else:
            path = os.path.join(REF_PYC, name + ".bin")
            loader = importlib.machinery.SourcelessFileLoader(name, path)
            spec = importlib.util.spec_from_loader(name, loader)
        mod = importlib.util.module_from_spec(spec)
        prev = sys.modules.get(name)
        sys.modules[name] = mod
        try:
            spec.loader.exec_module(mod)
        except BaseException:
            if prev is not None:
                sys.modules[name] = prev
            else:
                sys.modules.pop(name, None)
            raise
        _cache[name] = mod
        out[name] = mod
    return {n: out[n] for n in names}
