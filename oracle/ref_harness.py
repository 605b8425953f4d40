"""TEST INFRASTRUCTURE ONLY -- live-reference harness (runs only where /root/reference exists).

Imports the UNMODIFIED reference (DosepackAIR/MARL-DPP, mounted read-only at
/root/reference) so that golden vectors can be generated from it
(oracle/make_golden.py) and so that the CPU restatement in oracle/mdpp_oracle.c can be
pinned against it.  Nothing in the product package, in `-m gpu` tests, in smoke() or in
bench.py may import this module: /root/reference does not exist on the GPU box.

Recipe:
  * stub modules for tensorflow / keras / ordered_set (absent in this image; the env
    star-imports utils.py which imports them, reference src/utils/utils.py:1-9);
  * run from a writable scratch cwd (reference settings.py:4-6 creates ./models at import;
    bad-states path is cwd-relative, reference src/layout.py:6,247);
  * Display.__init__/update are no-ops (cv2 GUI, reference src/configuration.py:269-467).
"""
import contextlib
import copy
import io
import os
import shutil
import sys
import tempfile
import types

REFERENCE_ROOT = os.environ.get("MDPP_REFERENCE_ROOT", "/root/reference")

_loaded = None


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src"))


def _install_stubs():
    tf = types.ModuleType("tensorflow")
    tf.set_random_seed = lambda *a, **k: None
    tf.nn = types.SimpleNamespace(relu=None)
    keras = types.ModuleType("tensorflow.keras")
    tf.keras = keras
    py = types.ModuleType("tensorflow.python")
    fw = types.ModuleType("tensorflow.python.framework")
    ops = types.ModuleType("tensorflow.python.framework.ops")
    fw.ops = ops
    py.framework = fw
    tf.python = py
    for name, mod in [("tensorflow", tf), ("tensorflow.keras", keras), ("tensorflow.python", py),
                      ("tensorflow.python.framework", fw), ("tensorflow.python.framework.ops", ops)]:
        sys.modules.setdefault(name, mod)
    if "ordered_set" not in sys.modules:
        osm = types.ModuleType("ordered_set")

        class OrderedSet(list):
            def add(self, x):
                if x not in self:
                    self.append(x)
        osm.OrderedSet = OrderedSet
        sys.modules["ordered_set"] = osm


class Ref:
    """Handle on the imported reference modules plus a scratch cwd."""

    def __init__(self):
        self.scratch = tempfile.mkdtemp(prefix="mdpp_ref_")
        self._old_cwd = os.getcwd()
        os.chdir(self.scratch)
        _install_stubs()
        if REFERENCE_ROOT not in sys.path:
            sys.path.insert(0, REFERENCE_ROOT)
        with contextlib.redirect_stdout(io.StringIO()):
            import src.configuration as configuration
            import src.env_gym as env_gym
            import src.layout as layout
            import src.utils.utils as utils
            import src.algorithms.qlearning as qlearning
        configuration.Display.__init__ = lambda self, *a, **k: None
        configuration.Display.update = lambda *a, **k: -1
        self.configuration = configuration
        self.env_gym = env_gym
        self.layout = layout
        self.utils = utils
        self.qlearning = qlearning
        os.makedirs(os.path.join(self.scratch, "src/data/qvalue_files/layout-3"), exist_ok=True)

    # -- bad-states fixtures: "as-run" (off) vs "as-intended" (on), SURVEY fact 4 -------------
    def set_bad_states(self, on):
        dst = os.path.join(self.scratch, "data/bad_states_files/layout-3")
        if os.path.isdir(os.path.join(self.scratch, "data")):
            shutil.rmtree(os.path.join(self.scratch, "data"))
        if on:
            os.makedirs(dst, exist_ok=True)
            src = os.path.join(REFERENCE_ROOT, "src/data/bad_states_files/layout-3")
            for f in os.listdir(src):
                shutil.copy(os.path.join(src, f), os.path.join(dst, f))

    def make_configuration(self, block):
        with contextlib.redirect_stdout(io.StringIO()):
            return self.configuration.Configuration(copy.deepcopy(self.layout.layout_3[block]))

    def make_env(self, block, bad_states=False):
        self.set_bad_states(bad_states)
        cfg = self.make_configuration(block)
        with contextlib.redirect_stdout(io.StringIO()):
            env = self.env_gym.Environment(cfg)
        return env

    def make_rl(self, block, bad_states=False):
        self.set_bad_states(bad_states)
        cfg = self.make_configuration(block)
        with contextlib.redirect_stdout(io.StringIO()):
            rl = self.qlearning.RL(input_dim=80, output_dim=5, cars=5, sess=None, env_config=cfg)
        return rl, self.qlearning.env

    def close(self):
        os.chdir(self._old_cwd)
        shutil.rmtree(self.scratch, ignore_errors=True)


def load():
    global _loaded
    if _loaded is None:
        if not available():
            raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
        _loaded = Ref()
    return _loaded


@contextlib.contextmanager
def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        yield
