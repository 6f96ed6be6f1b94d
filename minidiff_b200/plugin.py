"""Backend plugin for the UNMODIFIED reference package (ahoynodnarb/minidiff).

    import sys; sys.argv += ["--backend", "minidiff_b200.plugin"]; import minidiff as md

The reference's loader (minidiff/backend/__init__.py:43-85) imports this module, takes the first
`Backend` subclass it finds in the module dict and copies every public attribute of that class
into `minidiff.backend`.  `ops/definitions.py`, `tensor.py`, `wrapping.py` and `topology.py` of
the reference then run unchanged on DeviceArray storage.  (Module rules per SURVEY 8b: import the
backend *module*, leave exactly one Backend subclass in this namespace.)

After `import minidiff`, `minidiff_b200.plugin.assert_live(md)` verifies the plugin was really
selected -- the reference silently falls back to NumPy if this import fails (finding 6).
"""
import minidiff.backend as backend  # the reference package; ImportError here if it is absent

from minidiff_b200.backend import functions as _F
from minidiff_b200.backend.device_array import DeviceArray as _DeviceArray

b200_backend = type(
    "b200_backend",
    (backend.Backend,),
    {name: (staticmethod(obj) if callable(obj) and not isinstance(obj, type) else obj)
     for name, obj in _F.TABLE.items()},
)


def assert_live(md_module) -> None:
    import minidiff.backend as live

    if live.tensor_class is not _DeviceArray:
        raise RuntimeError("the reference fell back to another backend; minidiff_b200 is not active")
