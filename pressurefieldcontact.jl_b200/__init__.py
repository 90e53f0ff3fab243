"""pfc-b200: B200-native contact-wrench evaluation for pressure-field contact.

The directory name contains a dot, so the package is imported under the alias ``pfc_b200``
(see ``pfc_b200.py`` at the repository root).
"""
from . import geometry  # noqa: F401
