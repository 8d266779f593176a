#!/usr/bin/env python
"""Drop-in for the reference's ``predict.py`` (same flags, same output files) on the B200 path.
The implementation lives in ``timed_design_b200/predict.py``."""
from timed_design_b200.predict import build_parser, load_dataset_and_predict, main  # noqa: F401

if __name__ == "__main__":
    main(build_parser().parse_args())
