"""Test-infrastructure shim: the reference imports `sound_field_analysis.sph.sphankel2`
at module level (core_deism.py:15, parallel_backends.py:18, core_deism_arg.py:15) but the
package is not installed in this image.  Only the one function is provided."""
