"""Hot-path subset of the reference's ``design_utils/utils.py`` backed by timed_design_b200."""
from timed_design_b200.frames import create_flat_dataset_map, load_batch  # noqa: F401
from timed_design_b200.postprocess import (  # noqa: F401
    compress_rotamer_predictions_to_20, convert_dataset_map_for_srb, extract_sequence_from_pred_matrix,
    get_pdb_keys_to_filter, get_rotamer_codec, load_datasetmap, save_consensus_probs, save_dict_to_fasta,
    save_outputs_to_file, standard_amino_acids)
