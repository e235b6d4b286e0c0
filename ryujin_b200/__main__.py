"""python -m ryujin_b200 <file.prm>: run one of the reference's parameter files on the B200 path"""
import sys

from .time_loop import main

sys.exit(main())
