"""Stand-in for the un-vendored `diffusers==0.24.0` package (reference environment.yaml:59), just wide enough for
the reference's own face_replace/models/unet_2d_condition/{unet,block}.py and attn_processors.py to import and run
in this container. Every leaf module re-exports the restatement in oracle/diffusers024.py. TEST INFRASTRUCTURE ONLY:
used by oracle/make_golden.py and tests that validate the oracle against the reference code (skipped when
/root/reference is absent)."""
__version__ = "0.24.0+oracle-shim"


class _NoPretrained:
    """Placeholder for the scheduler classes face_replace/models/model.py imports at module level; the oracle uses
    oracle.diffusers024.DDPMScheduler1Step instead (no hub access offline)."""

    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("no pretrained configs offline: use oracle.diffusers024.DDPMScheduler1Step")


class DDPMScheduler(_NoPretrained):
    pass


class EulerDiscreteScheduler(_NoPretrained):
    pass
