"""Minimal stand-in for the third-party `timm` package.

TEST INFRASTRUCTURE ONLY.  The reference (suous/RecNeXt) imports
`timm.layers` / `timm.models` at module import time (model/recnext.py:4-5,
model/recattn.py:4-5) but timm is not installed in this image and cannot be
(no network).  This shim provides just the five names the reference touches so
that the UNMODIFIED reference files can be imported by
`oracle/gen_golden.py` to produce golden vectors.  Nothing in the product
package imports it.
"""
from . import layers, models  # noqa: F401
