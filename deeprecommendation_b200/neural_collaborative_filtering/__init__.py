"""Host-side mirror of the reference's `neural_collaborative_filtering` package for the scoring hot path.

Same class names, constructor signatures, `forward` signatures, `state_dict` key names and collate contract as
/root/reference/src/neural_collaborative_filtering, so `train_model.py` / `evaluate_model.py`
switch over by changing the import root.  The arithmetic runs in libb200rec.so (CUDA, sm_100a).
"""
