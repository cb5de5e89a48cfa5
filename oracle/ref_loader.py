"""Locate the reference's *artefacts* (checkpoints) and, in the build container only, import the
reference's own Python modules to pin the oracle.  TEST INFRASTRUCTURE ONLY (see pram_oracle.py).

* ``/root/reference`` exists only in the build container.  ``stage_weights()`` (called from
  ``__graft_entry__.build()``) copies the two shipped checkpoints into ``oracle/_ref/weights/``,
  which is git-ignored (never enters history; licence CC BY-NC 4.0, reference LICENSE:1) but not
  gpurun-ignored, so the GPU box can run weight-dependent parity tests.  No reference *source* is
  ever copied.
* ``import_reference()`` puts ``/root/reference`` on ``sys.path`` and applies the one CPU shim the
  survey lists (nets.adagml.sink_algorithm hard-codes 'cuda', reference nets/adagml.py:45-48).
"""
from __future__ import annotations

import os
import shutil
import sys
from pathlib import Path
from typing import Optional

import numpy as np
import torch

REFERENCE_ROOT = Path(os.environ.get('PRAM_REFERENCE_ROOT', '/root/reference'))
HERE = Path(__file__).resolve().parent
STAGED = HERE / '_ref' / 'weights'

# inputs / weights shared with bench.py live outside oracle/ (benchdata.py); re-exported here for the tests
from benchdata import (SFD2_WEIGHT, GML_WEIGHT, weight_path, load_sfd2_state, load_gml_state, random_segnetvit_state,  # noqa: E402,F401
                       random_gml_state, random_sfd2_state, calibrated_adagml_state)


def reference_available() -> bool:
    return (REFERENCE_ROOT / 'nets' / 'sfd2.py').exists()


def stage_weights() -> None:
    if not reference_available():
        return
    STAGED.mkdir(parents=True, exist_ok=True)
    for name in (SFD2_WEIGHT, GML_WEIGHT):
        src, dst = REFERENCE_ROOT / 'weights' / name, STAGED / name
        if src.exists() and (not dst.exists() or dst.stat().st_size != src.stat().st_size):
            shutil.copyfile(src, dst)


def import_reference():
    """Returns a namespace of the reference's own modules (build container only)."""
    if not reference_available():
        raise RuntimeError('reference tree not mounted')
    if str(REFERENCE_ROOT) not in sys.path:
        sys.path.insert(0, str(REFERENCE_ROOT))
    import nets.sfd2 as ref_sfd2
    import nets.segnetvit as ref_segnetvit
    import nets.gml as ref_gml
    import nets.adagml as ref_adagml
    import nets.utils as ref_utils
    ref_adagml.sink_algorithm = ref_gml.sink_algorithm  # device-agnostic, identical maths
    from types import SimpleNamespace
    return SimpleNamespace(sfd2=ref_sfd2, segnetvit=ref_segnetvit, gml=ref_gml, adagml=ref_adagml,
                           utils=ref_utils)


