"""CPU oracle for the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything below this package.  The product
package ``deepfake_vit_b200`` never does (tests/test_boundary.py greps for it).

Contents
  efficientnet_pytorch/  restatement of the absent third-party dependency
                         efficientnet-pytorch==0.7.1 (pinned at task.ipynb:153)
  refmodel.py            restatement of the reference's own wrapper / attention /
                         head / loss code, used where /root/reference is not
                         mounted (the GPU box); proven bit-identical to the real
                         import on CPU by tests/test_oracle.py
  load_reference.py      imports the *real* reference modules from /root/reference
                         over the shim (this container only)
  make_golden.py         writes tests/golden/*.npz from the real import
  calibrate.py           the three weight sets of SURVEY.md section 7.1-2
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def ensure_shim_on_path():
    """Make ``import efficientnet_pytorch`` resolve to the restatement."""
    if _HERE not in sys.path:
        sys.path.insert(0, _HERE)
