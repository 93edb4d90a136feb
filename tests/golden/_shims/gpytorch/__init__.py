"""Test-only stand-in for the two gpytorch names the reference imports
(SURVEY.md Appendix C).  Never on the product path."""
from . import utils, distributions  # noqa: F401
