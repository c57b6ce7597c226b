"""`downstream` package shadowing the model modules of /root/reference/v2/downstream (zero-shot retrieval / recognition /
feature extraction use the pre-training towers forward-only; the evaluation scripts themselves are used from the checkout)."""
