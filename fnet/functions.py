"""Loader glue with the reference's names (fnet/functions.py:39-42)."""
import fnet.fnet_model


def load_model_from_path(opts, path_model_state, gpu_ids=0):
    model = fnet.fnet_model.Model(opts, gpu_ids=gpu_ids)
    model.load_state(path_model_state, gpu_ids=gpu_ids)
    return model
