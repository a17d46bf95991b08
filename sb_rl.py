#!/usr/bin/env python
"""Same command line as the reference's src/sb_rl.py, served by the B200 path:  python sb_rl.py -a PPO train -e Env01-v2"""
import pathlib
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
from balance_robot_b200.sb_rl import main  # noqa: E402

if __name__ == "__main__":
    main()
