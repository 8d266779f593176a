"""Import-compatibility shim: code written against the reference's ``design_utils`` package
(``from design_utils.utils import ...``, ``from design_utils.sampling_utils import ...``) resolves to
the B200 implementations of the hot-path functions.  Only the functions on the path are provided
(SURVEY.md section 8); plotting / SCWRL / AlphaFold helpers are out of scope."""
