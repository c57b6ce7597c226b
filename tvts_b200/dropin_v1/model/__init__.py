"""`model` package shadowing the hot-path modules of /root/reference/v1/model.  v1 and v2 use the same package
names, so the v1 drop-ins live in their own root: put tvts_b200/dropin_v1 (not tvts_b200/dropin) on PYTHONPATH for v1/train_dist_TVTS.py."""
