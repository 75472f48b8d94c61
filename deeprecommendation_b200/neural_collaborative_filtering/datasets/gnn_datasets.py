"""Graph datasets for GraphNCF (reference: datasets/gnn_datasets.py): samples are (user NODE id, item NODE id, target)."""
from __future__ import annotations

from .base import PointwiseDataset, RankingDataset


class _GraphMixin:
    def get_graph(self, device):
        return self.gcp.get_graph().to(device)


class GraphPointwiseDataset(_GraphMixin, PointwiseDataset):
    def __init__(self, file, graph_content_provider):
        super().__init__(file)
        self.gcp = graph_content_provider

    def __getitem__(self, item):
        userID, itemID, target = super().__getitem__(item)
        return self.gcp.get_user_nodeID(userID), self.gcp.get_item_nodeID(itemID), target

    @staticmethod
    def do_forward(model, batch, device, graph, *args):
        userIds, itemIds, y_batch = batch
        return model(graph.to(device), userIds.long().to(device), itemIds.long().to(device), device, *args), y_batch


class GraphRankingDataset(_GraphMixin, RankingDataset):
    def __init__(self, file, graph_content_provider):
        super().__init__(file)
        self.gcp = graph_content_provider

    def __getitem__(self, item):
        userID, item1ID, item2ID = super().__getitem__(item)
        return self.gcp.get_user_nodeID(userID), self.gcp.get_item_nodeID(item1ID), self.gcp.get_item_nodeID(item2ID)

    @staticmethod
    def do_forward(model, batch, device, graph, *args):
        userIds, item1Ids, item2Ids = batch
        g, u = graph.to(device), userIds.long().to(device)
        return (model(g, u, item1Ids.long().to(device), device, *args), model(g, u, item2Ids.long().to(device), device, *args))
