"""Builds development variants of the library (csrc/build.py tag=...) for A/B timing on the GPU box:

    python tools/variants.py            # builds every variant below into nerficg_b200/libnerf_b200.<tag>.so
    NERF_B200_LIB=nerficg_b200/libnerf_b200.<tag>.so python tools/kernel_timing.py
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from nerficg_b200.csrc import build as B  # noqa: E402

VARIANTS = {
    'nostore': ('NERF_EXP_NOSTORE',),            # diagnostic: chain kernels without their image stores (results are wrong downstream)
    'early': ('NERF_EXP_EARLY_HANDOFF',),        # operand hand-off before the slot barrier of the image store
    'split': ('NERF_EXP_SPLIT_STORE',),          # 4 x 16 KB stores instead of one 64 KB store
    'nohint': ('NERF_EXP_NOHINT',),              # stores without the L2 evict_first policy
    'earlysplit': ('NERF_EXP_EARLY_HANDOFF', 'NERF_EXP_SPLIT_STORE'),
    'lsu': ('NERF_EXP_LSU_STORE',),              # forward stash images leave through LSU store warps instead of TMA bulk stores
    'wrap': ('NERF_EXP_STORE_WRAP',),            # diagnostic: image stores hit a 16-tile window that stays in L2 (no HBM writes)
}

if __name__ == '__main__':
    import importlib
    names = sys.argv[1:] or list(VARIANTS)
    for tag in names:
        importlib.reload(B)
        print(tag, B.build(defines=VARIANTS[tag], tag=tag))
