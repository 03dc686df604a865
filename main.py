#!/usr/bin/env python
"""python main.py --config_path=configs/lqr_d5.json   (same CLI as the reference's main.py)"""
from deeppde_actorcritic_b200.main import run

if __name__ == "__main__":
    run()
