#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — builds a *runnable* copy of the reference outside the repo.

The reference checkout (/root/reference) does not import at HEAD:
flat layout vs. the `splendor/` package its imports expect, and an unfinished
function in SplendorLogicNumba.py (line 682-683) that is a SyntaxError.

This script copies the few files of the hot path into a scratch directory
(default /tmp/azg_ref, NEVER inside the repo: reference sources are not vendored)
and applies the mechanical patch list of SURVEY.md §8(c):

  P0  layout: root <- Game, MCTS, utils, NeuralNet ; splendor/ <- SplendorGame,
      SplendorLogic, SplendorLogicNumba, SplendorNNet, __init__
  P1  delete the stub `_valid_select_noble` (SplendorLogicNumba.py:682-683) and its
      call site (:262)
  P2  action_size(): 409 -> 406 (:35)   (last self-consistent action space, F4)
  P3  np.bool8 -> np.bool_ (:54)        (removed in NumPy 2)
  P4  delete the unfinished gold-allocation hook (:459-460)
  P5  colorama stub
  P6  pass (action 405) is a no-op that only bumps the ply counter (:284-285, F6)

Every hunk asserts that the text it replaces is present exactly once, so a changed
reference fails loudly instead of silently producing a different oracle.

  P7  (callers only) stubs for onnxruntime / coloredlogs / torchinfo, which Arena.py -> splendor/NNet.py ->
      GenericNNetWrapper.py and main.py import but this image does not have

Only oracle/refgen/gen_*.py and tests marked `ref` use the result.

`build(out, callers=True)` also copies the CALLERS of the hot path (Arena.py, Coach.py and what they import), unmodified, so that
tests can drive them over the reference's own Game / MCTS and over the B200 mirrors. The one place inside the repository a copy
may go is the git-ignored oracle/_ref/ (it travels to the GPU box with gpurun, never into history): `python
oracle/refgen/build_patched_ref.py --travel` builds oracle/_ref/pyref.
"""
import os
import shutil
import sys

REF = os.environ.get("AZG_REFERENCE", "/root/reference")
OUT = os.environ.get("AZG_REF_OUT", "/tmp/azg_ref")

ROOT_FILES = ["Game.py", "MCTS.py", "utils.py", "NeuralNet.py"]
PKG_FILES = ["SplendorGame.py", "SplendorLogic.py", "SplendorLogicNumba.py", "SplendorNNet.py"]
CALLER_ROOT_FILES = ["Arena.py", "Coach.py", "GenericNNetWrapper.py"]
CALLER_PKG_FILES = ["NNet.py"]
TRAVEL = os.path.realpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "_ref", "pyref"))


def _sub_once(text, old, new, tag):
    n = text.count(old)
    if n != 1:
        raise RuntimeError(f"patch {tag}: expected exactly 1 match, found {n}")
    return text.replace(old, new)


def patch_logic_numba(src):
    # P1: stub + call site
    src = _sub_once(src, "\tdef _valid_select_noble(player):\n\t\tif \n\n", "", "P1-stub")
    src = _sub_once(
        src,
        "\t\tresult[12+15+3+30+NUM_OF_EXCHANGE:-1] = self._valid_select_noble(player)\n",
        "",
        "P1-call",
    )
    # P2
    src = _sub_once(src, "\treturn 409 #+select noble pattern", "\treturn 406", "P2")
    # P3
    src = _sub_once(src, "dtype=np.bool8)", "dtype=np.bool_)", "P3")
    # P4
    src = _sub_once(
        src,
        "\t\tif (num_gold := self.players_gems[player][idx_gold]) > 0:\n"
        "\t\t\tgalloc_pattern = self._calc_gold_alloc(card0, player, num_gold)",
        "\t\tpass",
        "P4",
    )
    # the hook's callee returns None (untypable in numba) - drop it as well
    start = src.index("\tdef _calc_gold_alloc(self, card, player, num_gold):")
    end = src.index("\tdef _buy_card(self, card0, card1, player):")
    src = src[:start] + src[end:]
    # P6: pass = no-op
    src = _sub_once(
        src,
        "\t\telse: #365-404\n\t\t\tself._give_and_get_gems(move-60, player)\n",
        "\t\telif move < 405: #365-404\n\t\t\tself._give_and_get_gems(move-60, player)\n",
        "P6",
    )
    return src


COLORAMA_STUB = '''"""stub (P5): console colours are UI-only"""
class _Blank:
    def __getattr__(self, name):
        return ""
Style = _Blank(); Fore = _Blank(); Back = _Blank()
'''


STUBS = {   # P7
    "onnxruntime.py": '"""stub (P7): the ONNX path is not selected at HEAD (GenericNNetWrapper.py:25-29)"""\n',
    "coloredlogs.py": '"""stub (P7)"""\ndef install(*a, **k):\n    pass\n',
    "torchinfo.py": '"""stub (P7)"""\ndef summary(*a, **k):\n    return None\n',
}


def build(out=OUT, ref=REF, callers=False):
    if not os.path.isdir(ref):
        raise FileNotFoundError(ref)
    real_out = os.path.realpath(out)
    repo = os.path.realpath(os.path.join(os.path.dirname(__file__), "..", ".."))
    if real_out.startswith(repo + os.sep) and real_out != TRAVEL:
        raise RuntimeError("refusing to copy reference sources into the repo (only the git-ignored oracle/_ref/pyref may hold a copy)")
    if os.path.isdir(out):
        shutil.rmtree(out)
    os.makedirs(os.path.join(out, "splendor"))
    for f in ROOT_FILES:
        shutil.copy(os.path.join(ref, f), os.path.join(out, f))
    for f in PKG_FILES:
        with open(os.path.join(ref, f)) as fh:
            src = fh.read()
        if f == "SplendorLogicNumba.py":
            src = patch_logic_numba(src)
        with open(os.path.join(out, "splendor", f), "w") as fh:
            fh.write(src)
    open(os.path.join(out, "splendor", "__init__.py"), "w").close()
    with open(os.path.join(out, "colorama.py"), "w") as fh:
        fh.write(COLORAMA_STUB)
    if callers:
        for f in CALLER_ROOT_FILES:
            shutil.copy(os.path.join(ref, f), os.path.join(out, f))
        for f in CALLER_PKG_FILES:
            shutil.copy(os.path.join(ref, f), os.path.join(out, "splendor", f))
        for name, text in STUBS.items():
            with open(os.path.join(out, name), "w") as fh:
                fh.write(text)
    return out


def import_ref(out=OUT, callers=False):
    """Build (if needed) and put the patched copy on sys.path."""
    if not os.path.isfile(os.path.join(out, "splendor", "SplendorLogicNumba.py")) or (callers and not os.path.isfile(os.path.join(out, "Coach.py"))):
        build(out, callers=callers)
    if out not in sys.path:
        sys.path.insert(0, out)
    return out


def find_ref(callers=False):
    """the patched copy that is available: built from /root/reference when that exists (build container), else the
    travelling copy oracle/_ref/pyref (GPU box), else None"""
    if os.path.isdir(REF):
        return import_ref(callers=callers)
    if os.path.isfile(os.path.join(TRAVEL, "Coach.py" if callers else os.path.join("splendor", "SplendorLogicNumba.py"))):
        if TRAVEL not in sys.path:
            sys.path.insert(0, TRAVEL)
        return TRAVEL
    return None


if __name__ == "__main__":
    if "--travel" in sys.argv:
        print(build(TRAVEL, callers=True))
    else:
        print(build(callers="--callers" in sys.argv))
