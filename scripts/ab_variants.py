"""Development aid: build A/B variants of libb200_tetris_mcts.so (compile-time switches of the same source) and print the gpurun job
that times them back to back in ONE GPU session (scripts/exp_growth.py per variant).  Results of round 1: profiles/ab_variants_r1.txt.

  python scripts/ab_variants.py base:-DB200_WARM_EXPAND=0,-DB200_WARM_SELECT=0 fused:-DB200_FUSED_BACKUP=1 > scripts/_ab_run.sh
  gpurun --timeout 900 -- 'bash scripts/_ab_run.sh'

Switches (defaults in the sources): B200_WARM_EXPAND, B200_WARM_SELECT, B200_L2_HOT_LEVELS, B200_L2_STREAM_ACT3, B200_FUSED_BACKUP,
B200_BACKUP_MINB, B200_PLAY_UNIFIED.  The variants land in build/variants/ (git-ignored, travels to the GPU box); a variant is
loaded instead of the product library when B200_TETRIS_LIB names it (tetris_mcts_b200/_lib.py)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tetris_mcts_b200 import build as B  # noqa: E402


def main():
    specs = [a.split(":", 1) for a in sys.argv[1:]]
    out_dir = os.path.join(ROOT, "build", "variants")
    os.makedirs(out_dir, exist_ok=True)
    with ThreadPoolExecutor(4) as ex:
        list(ex.map(lambda s: B.build(out=os.path.join(out_dir, "lib_%s.so" % s[0]), defines=[d for d in s[1].split(",") if d]), specs))
    B.build()
    names = [s[0] for s in specs] + ["default"]
    print("cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out")
    print("for v in %s; do" % " ".join(names))
    print('  echo "=== variant $v"')
    print("  if [ $v = default ]; then unset B200_TETRIS_LIB; else export B200_TETRIS_LIB=$GRAFT_REPO_ROOT/build/variants/lib_$v.so; fi")
    print("  timeout 300 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | grep -E \"^move|^\\{'select|k_select\" | cut -c1-400")
    print("  NO_TIMING=1 GC_HEADROOM=2560 timeout 300 python scripts/exp_growth.py 16384 16384 500 4 net_tc 2>&1 | grep -E \"^move\" | cut -c1-200")
    print("done 2>&1 | tee gpurun_out/exp_variants.log")


if __name__ == "__main__":
    main()
