"""Run the reference's UNMODIFIED driver script (train_sr_dr.py / train_sr.py, from oracle/_ref) with amid_b200 as the
drop-in `model_seq` module: the driver's own argparse, datasets, DataLoader workers, samplers, train()
and test() loops, torch.optim.Adam pair and logging run untouched; only `from model_seq import *` resolves to
amid_b200/dropin/model_seq.py.

    python tools/run_reference_driver.py train_sr_dr.py --epoch 1 --model sasrec --isItC True --ts2 0.4 \
        -ds amazon -dm cloth_sport --overlap_ratio 0.75 --neg_nums 199 --lr2 0.01 --dr_e_w 0.01

The only patch is shim 1 of SURVEY.md 8c (random.sample on a set, a Python >= 3.11 incompatibility of the reference's
sampler), installed before the script starts so that the forked DataLoader workers inherit it.
"""
import os
import random
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
_orig = random.sample


def _sample(pop, k, **kw):
    if isinstance(pop, (set, frozenset)):
        pop = tuple(pop)
    return _orig(pop, k, **kw)


def main():
    if not os.path.exists(os.path.join(REF, "train_sr_dr.py")):
        raise SystemExit("oracle/_ref is missing: run `python oracle/build_ref.py` where /root/reference exists")
    script = sys.argv[1]
    random.sample = _sample
    # import order: the drop-in model_seq first, then the reference's own dataset_seq / utils
    sys.path[:0] = [os.path.join(ROOT, "amid_b200", "dropin"), ROOT, REF]
    os.chdir(REF)                                      # the drivers use relative CSV / log paths (train_sr_dr.py:636-642)
    os.makedirs("model", exist_ok=True)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(os.path.join(REF, script), run_name="__main__")


if __name__ == "__main__":
    main()
