"""Two-function stand-in for the `klepto` package (absent from this image).

TEST INFRASTRUCTURE ONLY.  mystic hard-imports `klepto.isvalid/validate` at
mystic/abstract_solver.py:80 and uses them only in SetObjective (:770-774) as an
argument-count sanity check.  This shim lets the UNMODIFIED reference under
/root/reference be imported in the build container so that golden vectors can be
recorded from it (oracle/record_reference.py).  It never travels into the product.
"""


def isvalid(func, *args, **kwds):
    return True


def validate(func, *args, **kwds):
    return None
