"""stand-in for machines WITHOUT pytorch3d (do not put this directory on sys.path when the real package is installed)"""
