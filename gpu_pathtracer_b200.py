"""Import shim: the package directory is named `gpu-pathtracer_b200` (not a valid Python identifier), so
`import gpu_pathtracer_b200` loads it from there."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu-pathtracer_b200")
_spec = importlib.util.spec_from_file_location("gpu_pathtracer_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["gpu_pathtracer_b200"] = _mod
_spec.loader.exec_module(_mod)

if __name__ == "__main__":                      # python gpu_pathtracer_b200.py scene.json --spp N --png out.png
    from gpu_pathtracer_b200.cli import main
    sys.exit(main())
