"""The reference's ``design_utils/sampling_utils.py`` API backed by the GPU sampler."""
from timed_design_b200.sampling_utils import (  # noqa: F401
    apply_temp_to_probs, random_choice_prob_index, sample_from_sequences, sample_with_multiprocessing,
    save_as, set_seed)
