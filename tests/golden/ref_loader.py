"""Import the UNMODIFIED reference (/root/reference) with import-time shims only.

Test infrastructure.  Used by ``gen_golden.py`` (to produce the committed
fixtures) and by the optional live cross-check tests, which skip when
``/root/reference`` is absent (it does not exist on the GPU box).

The shims do not touch reference arithmetic:
  * ``termcolor`` is not installed         -> stub ``colored`` that returns the text
  * ``matplotlib.pyplot`` is not installed -> no-op stub (only PNG dumps use it,
    new_quantity_op.py:341-355)
  * ``time.clock`` was removed in py3.8    -> ``time.perf_counter``
    (pytorch_quantizer.py:293,295,420,425)
  * ``yaml.load`` needs a Loader in PyYAML 6 -> default to SafeLoader
    (pytorch_quantizer.py:28,32; reconstruction.py:105)

No reference source is copied: the python modules are imported from where they
lie; only the two yml files the drivers read cwd-relatively are re-created in a
scratch dir under /tmp so the drivers can write ./workdir there.
"""
import contextlib
import importlib
import importlib.util
import os
import sys
import tempfile
import time
import types

def _default_root():
    """/root/reference in the build container; the staged byte copy baseline/_ref (stage_ref.py) elsewhere."""
    if os.path.isdir("/root/reference/quantity"):
        return "/root/reference"
    return os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "baseline", "_ref")


REF_ROOT = os.environ.get("PQ_REFERENCE_ROOT") or _default_root()
REF_QUANTITY = os.path.join(REF_ROOT, "quantity")


def available():
    return os.path.isdir(os.path.join(REF_QUANTITY, "common", "quantity"))


def install_shims():
    if "termcolor" not in sys.modules:
        m = types.ModuleType("termcolor")
        m.colored = lambda s, *a, **k: s
        sys.modules["termcolor"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")

        class _Fig:
            def savefig(self, *a, **k):
                pass

        plt.figure = lambda *a, **k: _Fig()
        for name in ("grid", "title", "xlabel", "ylabel", "hist", "close"):
            setattr(plt, name, lambda *a, **k: None)
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if not hasattr(time, "clock"):
        time.clock = time.perf_counter
    import yaml

    if not getattr(yaml.load, "_pq_shim", False):
        _orig = yaml.load

        def _load(stream, Loader=None, **kw):
            return _orig(stream, Loader=Loader or yaml.SafeLoader, **kw)

        _load._pq_shim = True
        yaml.load = _load


def load_l2(name):
    """Load one self-contained reference L2 file (quantity/common/quantity/<name>.py)
    under a private module name so it can coexist with this repo's ``common``."""
    install_shims()
    mod_name = "pq_ref_" + name
    if mod_name in sys.modules:
        return sys.modules[mod_name]
    path = os.path.join(REF_QUANTITY, "common", "quantity", name + ".py")
    spec = importlib.util.spec_from_file_location(mod_name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[mod_name] = mod  # multiprocessing pickles functions by module name
    spec.loader.exec_module(mod)
    return mod


CONFIGS_YML = """OUTPUT:
    WORK_DIR: ./workdir
    WEIGHT_BIT_TABLE: ./workdir/weight.table
    FEAT_BIT_TABLE: ./workdir/feat.table
    WEIGHT_DIR: ./workdir/weight
    BIAS_DIR: ./workdir/bias
    FINAL_WEIGHT_DIR: ./workdir/new_weight
    FINAL_BIAS_DIR: ./workdir/new_bias
SETTINGS:
    GPU: 0
    WORKER_NUM: {worker_num}
    INTERVAL_NUM: 2048
    STATISTIC: 1
    MAX_CALI_IMG_NUM: {max_cali}
    MAX_SHIFT: 12
    SUPPORT_DILATION: False
    MERGE_FREEZEBN: False
    CARE_OP_TYPE: ['Conv2d', 'Linear', 'Eltwise', 'Concat']
    ALL_OP_TYPE: ['Conv2d', 'Linear', 'Eltwise', 'Concat', 'MaxPool2d', 'ReLU', 'UpsamplingNearest2d', 'View', 'AvgPool2d']
    ALLOW_SAME_TID_OP_TYPE: ['ReLU', 'UpsamplingNearest2d', 'View']
    MERGE_OP_YTPE: ['Eltwise', 'Concat']
"""

USER_CONFIGS_YML = """PATH:
    DATA_PATH: ./none
    MODEL_NET_PATH: ./none.py
    MODEL_PATH: ./none.pth
    QUANTITY_MODEL_PATH: ./workdir/quantity_model.pth
MODEL:
    INPUT_SHAPE: {input_shape}
PRE_PROCESS:
    IMG: 1
    IMG_SET:
      MEAN: 128
      RESIZE: 32,32
      SCALE: 0.0075
SETTINGS:
    DEVICE: {device}
    GPU: 0
"""


def stage_workdir(input_shape, max_cali, worker_num=4, device="cpu", root=None):
    """Create <root>/tools/configs.yml and <root>/test/user_configs.yml (the keys of
    quantity/tools/configs.yml:12-33 and quantity/test/user_configs.yml:1-25) and
    return <root>/test, the cwd the drivers must run in."""
    root = root or tempfile.mkdtemp(prefix="pq_ref_stage_")
    os.makedirs(os.path.join(root, "tools"), exist_ok=True)
    os.makedirs(os.path.join(root, "test"), exist_ok=True)
    with open(os.path.join(root, "tools", "configs.yml"), "w") as f:
        f.write(CONFIGS_YML.format(worker_num=worker_num, max_cali=max_cali))
    with open(os.path.join(root, "test", "user_configs.yml"), "w") as f:
        f.write(USER_CONFIGS_YML.format(
            input_shape=",".join(str(int(v)) for v in input_shape), device=device))
    return os.path.join(root, "test")


@contextlib.contextmanager
def reference_tools(input_shape, max_cali, worker_num=4, extra_sys_path=(), device="cpu", foreign_common=False):
    """Context: cwd = staged test dir, sys.path[0] = reference quantity/, yields the
    reference ``tools`` package.  Must not be used in a process that already
    imported this repo's ``common``/``tools`` (same top-level names by design)."""
    install_shims()
    for k in list(sys.modules):
        if foreign_common and (k == "common" or k.startswith("common.")):
            continue            # boundary proof: the reference's tools/ on top of the drop-in common.quantity
        if k == "common" or k.startswith("common.") or k == "tools" or k.startswith("tools."):
            f = getattr(sys.modules[k], "__file__", "") or ""
            assert f.startswith(REF_ROOT) or not f, (
                "this process already imported a non-reference '%s' (%s)" % (k, f))
    cwd0 = os.getcwd()
    test_dir = stage_workdir(input_shape, max_cali, worker_num, device=device)
    paths = [REF_QUANTITY] + list(extra_sys_path)
    for p in reversed(paths):
        sys.path.insert(0, p)
    os.chdir(test_dir)
    try:
        tools = importlib.import_module("tools")
        yield tools, test_dir
    finally:
        os.chdir(cwd0)
        for p in paths:
            sys.path.remove(p)
