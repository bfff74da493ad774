"""Drop-in namesake of the reference's ``miscc`` package (only the hot-path module ``losses`` and the
``cfg`` object it reads)."""
