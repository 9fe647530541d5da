"""The software pipeline of the sweep kernels, checked in the compiled code (no GPU needed): the loads a level issues
for the NEXT level must not be waited for before the level's own work is done.

ptxas tracks outstanding loads with six counting scoreboards and puts nearly every global load of these kernels on one
of them, so a single misplaced wait switches the pipeline off: in round 2 the first use of the slices at the top of
the finite-width level carried such a wait and every level stalled for its own prefetches (13.5 % of C4's stall
samples, DESIGN.md "Scoreboards").  This test decodes the scheduling control words of `cuobjdump -sass`
(scripts/sass_ctrl.py) and requires, for every production 2^popcount kernel of 32 lanes, that the first wait on the
scoreboard of the level's prefetch loads comes at least 60 instructions behind the level's votes."""
import os
import shutil
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'scripts'))

CUOBJDUMP = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'


def _listing(obj, name):
    from sass_ctrl import decode
    out = subprocess.run([CUOBJDUMP, '-sass', '-fun', name, obj], capture_output=True, text=True)
    if out.returncode != 0 or 'Function' not in out.stdout:
        pytest.skip('cuobjdump could not list ' + name)
    return decode(out.stdout.splitlines())


# (object, mangled kernel name): sa_sweep_kernel<32, WPL, FINITE, RngPhilox<32>, DIM2 = true, 28, false, false>
KERNELS = [
    ('tnb_inst_32_1.o', '_ZN3tnb15sa_sweep_kernelILi32ELi1ELb1ENS_9RngPhiloxILi32EEELb1ELi28ELb0ELb0EEEvNS_6ParamsE'),  # C4
    ('tnb_inst_32_1.o', '_ZN3tnb15sa_sweep_kernelILi32ELi1ELb0ENS_9RngPhiloxILi32EEELb1ELi28ELb0ELb0EEEvNS_6ParamsE'),  # C2, C3
    ('tnb_inst_32_2.o', '_ZN3tnb15sa_sweep_kernelILi32ELi2ELb0ENS_9RngPhiloxILi32EEELb1ELi28ELb0ELb0EEEvNS_6ParamsE'),  # C5
]


@pytest.mark.parametrize('obj,name', KERNELS)
def test_level_prefetches_are_not_waited_for_at_the_level_top(obj, name):
    path = os.path.join(ROOT, 'tnco_b200', 'csrc', 'build', obj)
    if not os.path.exists(path) or not os.path.exists(CUOBJDUMP):
        pytest.skip('no compiled kernels / cuobjdump here (run __graft_entry__.build() first)')
    ins = _listing(path, name)
    # level tops: the pair of votes on "does a child of B share an index with C" -- two VOTE.ANY on predicates within
    # four instructions of each other (the unconstrained kernels hold two copies of the level)
    votes = [i for i, x in enumerate(ins) if x[1].startswith('VOTE.ANY P')]
    tops = [i for i in votes if any(0 < j - i <= 4 for j in votes)]
    assert tops, 'no level found in ' + name
    checked = 0
    for top in tops:
        # the prefetch loads of this level: global loads in the 45 instructions in front of the votes
        loads = [x for x in ins[max(0, top - 45):top] if x[1].lstrip('@!P0123456 ').startswith('LDG') and x[2] is not None]
        if len(loads) < 3:
            continue  # (not a level: the re-slicer votes too)
        sbs = {x[2] for x in loads}
        first_wait = next((k for k in range(top, min(len(ins), top + 400)) if ins[k][4] & sbs), None)
        assert first_wait is not None
        assert first_wait - top >= 60, (
            f'{name}: instruction {ins[first_wait][0]:#x} "{ins[first_wait][1]}" waits for the level\'s prefetch loads '
            f'{first_wait - top} instructions behind the votes -- the software pipeline is off (see DESIGN.md, Scoreboards)')
        checked += 1
    assert checked >= 1
