#!/usr/bin/env python
"""Drop-in for the reference's ``sample.py`` (same flags, same output files) on the B200 path.
The implementation lives in ``timed_design_b200/sample.py``."""
from timed_design_b200.sample import build_parser, main_sample  # noqa: F401

if __name__ == "__main__":
    main_sample(build_parser().parse_args())
