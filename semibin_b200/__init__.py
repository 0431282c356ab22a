"""semibin_b200 -- B200-native clustering front end for SemiBin (kNN graph,
abundance edge weights, eps-sweep DBSCAN).  See DESIGN.md."""
__version__ = '0.1.0'
