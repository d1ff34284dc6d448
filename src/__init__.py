"""Reference-named package: the drivers import `src.PGAS`, `src.Filtering`, ...."""
