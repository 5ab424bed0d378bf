"""The tcgen05.mma issue loops must stay free of the compiler's per-lane serialisation loop.

tcgen05.mma / tcgen05.commit take uniform-register operands.  Inside an `if (lane == 0)` region ptxas cannot prove that one lane is
active and wraps EVERY such instruction in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (100-160 cycles of dependent issue per MMA:
profiles/r02_pool_phase_cycles.txt); under `elect.sync` (umma_ptx.cuh: elect_one) they issue back to back.  This test disassembles
the built objects and fails if a UTCHMMA is followed by that loop's back-branch again."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, 'pytorch_graphsage_b200', 'build')
KERNEL_OBJECTS = ['linear_ws_umma', 'linear_pool_ws_umma', 'attention_umma', 'wgrad_umma', 'gather_mean_project_umma', 'linear_pool_umma']


@pytest.mark.parametrize('name', KERNEL_OBJECTS)
def test_mma_issue_is_not_wrapped_in_a_lane_loop(name):
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    obj = os.path.join(OBJ, name + '.o')
    if not os.path.exists(cuobjdump) or not os.path.exists(obj):
        pytest.skip('needs cuobjdump and the objects of an in-tree build (python -m pytorch_graphsage_b200.build)')
    sass = subprocess.run([cuobjdump, '-sass', obj], capture_output=True, text=True, check=True).stdout
    lines = [l for l in sass.splitlines() if '/*' in l and ';' in l and not l.strip().startswith('/* 0x')]
    mma = [i for i, l in enumerate(lines) if 'UTCHMMA' in l]
    assert mma, 'no tcgen05.mma in %s: wrong object?' % name
    wrapped = [i for i in mma if any('BRA.U.ANY' in l for l in lines[i + 1:i + 4])]
    assert not wrapped, '%d of %d UTCHMMA in %s sit in an ELECT / BRA.U.ANY loop: issue them under elect_one(), not lane == 0' % (len(wrapped), len(mma), name)
