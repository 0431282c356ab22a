"""Bind an installed SemiBin to the GPU path (the seam SURVEY.md section 8(b)
describes: "replace these call sites / monkey-patch these names").

    import semibin_b200.patch; semibin_b200.patch.install()

After install(), `SemiBin.cluster.run_embed_infomap` is the fused GPU version, `SemiBin.cluster.recluster_bins` runs
all its per-bin KMeans problems as one batched launch,
`SemiBin.long_read_cluster.cluster_long_read` is the GPU version (kNN graph,
12-eps DBSCAN sweep and the get_best_bin extraction loop on the device; main.py
imports the name at call time, :1170) and `SemiBin.long_read_cluster`'s
`kneighbors_graph` / `dbscan` / `get_best_bin`-free call sites resolve to the
libsbk-backed drop-ins; `uninstall()` restores the originals.
"""
_saved = {}


def install():
    import SemiBin.cluster as rc
    import SemiBin.long_read_cluster as rl
    from . import cluster, long_read, neighbors, recluster
    if _saved:
        return
    _saved.update(run_embed_infomap=rc.run_embed_infomap, kneighbors_graph=rl.kneighbors_graph, dbscan=rl.dbscan,
                  cluster_long_read=rl.cluster_long_read, recluster_bins=rc.recluster_bins)
    import functools
    rc.run_embed_infomap = cluster.run_embed_infomap
    rc.recluster_bins = recluster.recluster_bins            # cluster() calls it by its module-global name (cluster.py:304)
    # the graph stays on the device for the 12 dbscan() calls that follow it (long_read_cluster.py:108-120)
    rl.kneighbors_graph = functools.partial(neighbors.kneighbors_graph, keep_device=True)
    cache = long_read._SweepCache()
    neighbors._graph_listeners.append(cache.reset)          # a new graph invalidates the sweep cache
    _saved['listener'] = cache.reset
    rl.dbscan = cache
    rl.cluster_long_read = long_read.cluster_long_read


def uninstall():
    if not _saved:
        return
    import SemiBin.cluster as rc
    import SemiBin.long_read_cluster as rl
    rc.run_embed_infomap = _saved['run_embed_infomap']
    rc.recluster_bins = _saved['recluster_bins']
    rl.kneighbors_graph = _saved['kneighbors_graph']
    rl.dbscan = _saved['dbscan']
    rl.cluster_long_read = _saved['cluster_long_read']
    from . import neighbors
    if _saved.get('listener') in neighbors._graph_listeners:
        neighbors._graph_listeners.remove(_saved['listener'])
    _saved.clear()
