"""Stand-in for the un-vendored `diffusers==0.24.0` package (reference environment.yaml:59), just wide enough for
the reference's own face_replace/models/unet_2d_condition/{unet,block}.py and attn_processors.py to import and run
in this container. Every leaf module re-exports the restatement in oracle/diffusers024.py. TEST INFRASTRUCTURE ONLY:
used by oracle/make_golden.py and tests that validate the oracle against the reference code (skipped when
/root/reference is absent)."""
__version__ = "0.24.0+oracle-shim"
