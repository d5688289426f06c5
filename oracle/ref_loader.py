"""Load the reference's own BPMF/clib.py against oracle/_ref/lib/libc.so (TEST INFRASTRUCTURE).

``import BPMF`` fails in this image (BPMF/__init__.py:8-14 needs obspy/h5py/matplotlib), so a stub
package is registered and only ``config.py`` and ``clib.py`` are executed, with ``cfg.PACKAGE``
pointed at a directory holding ``lib/libc.so`` (clib.py:12-15).  Works only where /root/reference
exists (this container); the GPU box uses the committed fixtures in tests/golden/ instead.
"""
import importlib.util
import os
import sys
import types

REF = "/root/reference/BPMF"
HERE = os.path.dirname(os.path.abspath(__file__))


def available():
    return os.path.exists(os.path.join(REF, "clib.py"))


def load_reference_clib(package_dir=None):
    """Return the reference ``BPMF.clib`` module bound to ``<package_dir>/lib/libc.so``."""
    from . import build as _build

    if package_dir is None:
        _build.build_ref()
        package_dir = os.path.join(HERE, "_ref")
    name = "BPMF_ref_%d" % abs(hash(package_dir))
    pkg = types.ModuleType(name)
    pkg.__path__ = [REF]
    sys.modules[name] = pkg
    cwd = os.getcwd()
    try:
        os.chdir(HERE)  # config.py reads ./BPMF_parameters.cfg from CWD; none here -> defaults
        spec = importlib.util.spec_from_file_location(name + ".config", os.path.join(REF, "config.py"))
        config = importlib.util.module_from_spec(spec)
        sys.modules[name + ".config"] = config
        try:
            spec.loader.exec_module(config)
        except Exception:
            # config.py may fail without a parameter file; provide the one attribute clib needs
            class _Cfg:
                pass
            config.cfg = _Cfg()
        config.cfg.PACKAGE = package_dir
        spec = importlib.util.spec_from_file_location(name + ".clib", os.path.join(REF, "clib.py"))
        clib = importlib.util.module_from_spec(spec)
        clib.__package__ = name
        sys.modules[name + ".clib"] = clib
        spec.loader.exec_module(clib)
    finally:
        os.chdir(cwd)
    if not clib.cpu_loaded:
        raise RuntimeError("reference clib.py could not load %s/lib/libc.so" % package_dir)
    return clib
