"""The continuous crawl loop of ``infomesh start``: seed -> schedule -> crawl -> index (or submit to indexer peers) ->
publish -> credit; RSS polling, the priority recrawl queue, governor back-pressure, disk guard, idle re-seeding and
hourly FTS optimisation (reference infomesh/crawler/crawl_loop.py:37-501).

Structure differs from the reference's single long function: each duty is a small coroutine driven by a table of
(interval, action) pairs, and seed rediscovery is shared between start-up and idle re-seeding.
"""
from __future__ import annotations

import asyncio
import contextlib
import time
from typing import Any

from infomesh_b200.crawler.parser import extract_links
from infomesh_b200.crawler.seeds import CATEGORIES, load_seeds
from infomesh_b200.resources.preflight import is_disk_critically_low
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

PRIORITY_BATCH = 5
IDLE_RESTART_SECONDS = 10.0
FTS_OPTIMIZE_INTERVAL = 3600.0
PRIORITY_CHECK_INTERVAL = 2.0
GOVERNOR_CHECK_INTERVAL = 5.0
DISK_CHECK_INTERVAL = 60.0
GPU_REBUILD_PENDING = 256            # rebuild the HBM mirror after this many newly indexed documents


def _credit(ctx: Any, note: str) -> None:
    if getattr(ctx, "ledger", None) is None:
        return
    try:
        from infomesh_b200.credits.ledger import ActionType

        ctx.ledger.record_action(ActionType.CRAWL, quantity=1.0, note=note, key_pair=ctx.key_pair)
    except Exception:  # noqa: BLE001 — accounting must never stop the crawl
        logger.debug("credit_record_failed", note=note[:60])


async def _index_and_publish(ctx: Any, result: Any) -> None:
    from infomesh_b200.services import index_document, publish_document_to_network

    doc_id = index_document(result.page, ctx.store, ctx.vector_store, js_required=result.js_required)
    await publish_document_to_network(result.page, doc_id, p2p_node=ctx.p2p_node, distributed_index=ctx.distributed_index)
    gi = getattr(ctx, "gpu_index", None)
    if gi is not None and doc_id is not None:
        gi.note_added()
        if gi._pending >= GPU_REBUILD_PENDING:          # noqa: SLF001
            await asyncio.to_thread(getattr(gi, "refresh", gi.rebuild))     # incremental append when the index supports it


async def feed_poll_loop(ctx: Any) -> None:
    """Poll due RSS/Atom feeds; every new entry URL goes to the priority recrawl queue."""
    monitor, queue = getattr(ctx, "feed_monitor", None), getattr(ctx, "priority_queue", None)
    if monitor is None or queue is None or ctx.worker is None:
        return
    from infomesh_b200.crawler.freshness import RecrawlTrigger

    while True:
        due = monitor.get_due_feeds()
        if not due:
            await asyncio.sleep(10)
            continue
        for feed in due:
            try:
                client = await ctx.worker.get_http_client()
                resp = await client.get(feed.url, timeout=30.0)
                if resp.status_code < 400:
                    for url in monitor.process_feed_response(feed.url, resp.text).new_urls:
                        queue.enqueue(url, RecrawlTrigger.RSS_UPDATE, source_feed=feed.url)
                    continue
                logger.warning("feed_poll_http_error", url=feed.url, status=resp.status_code)
            except Exception as exc:  # noqa: BLE001 — network errors of any client library
                logger.debug("feed_poll_failed", url=feed.url, error=str(exc))
            feed.error_count += 1
            feed.last_poll_at = time.time()
        await asyncio.sleep(5)


async def _process_priority_queue(ctx: Any, _logger: Any = None) -> int:
    queue = getattr(ctx, "priority_queue", None)
    if queue is None or queue.size == 0 or ctx.worker is None:
        return 0
    done = 0
    for _ in range(PRIORITY_BATCH):
        item = queue.dequeue()
        if item is None:
            break
        try:
            res = await ctx.worker.crawl_url(item.url, depth=0)
            if res.success and res.page:
                await _index_and_publish(ctx, res)
                done += 1
                _credit(ctx, f"priority:{item.trigger}:{item.url[:100]}")
            logger.info("priority_crawl", url=item.url, trigger=str(item.trigger), success=res.success)
        except Exception:  # noqa: BLE001
            logger.exception("priority_crawl_failed")
    return done


async def _apply_governor_backpressure(ctx: Any, _logger: Any = None) -> bool:
    """True -> skip this iteration (crawl paused)."""
    gov = getattr(ctx, "governor", None)
    if gov is None:
        return False
    st = gov.check_and_adjust()
    if gov.should_pause_crawl:
        logger.warning("governor_pause", level=st.degrade_level.name, cpu=f"{st.cpu_percent:.0f}%", mem=f"{st.memory_percent:.0f}%")
        await asyncio.sleep(10)
        return True
    if gov.should_throttle_crawl:
        await asyncio.sleep(max(0.1, (1.0 - st.throttle_factor) * 2.0))
    return False


async def _enqueue_seed(ctx: Any, url: str) -> tuple[int, int]:
    """Unseen seed -> queue it; seen seed -> refetch and queue its unseen links.  -> (new, rediscovered)."""
    if not ctx.dedup.is_url_seen(url):
        return (1, 0) if await ctx.scheduler.add_url(url, depth=0) else (0, 0)
    found = 0
    try:
        client = await ctx.worker.get_http_client()
        resp = await client.get(url, timeout=30.0)
        if resp.status_code < 400:
            for link in extract_links(resp.text, url):
                if not ctx.dedup.is_url_seen(link) and await ctx.scheduler.add_url(link, depth=1):
                    found += 1
    except Exception as exc:  # noqa: BLE001
        logger.debug("seed_rediscovery_failed", url=url, error=str(exc))
    return 0, found


async def _reseed_queue(ctx: Any, _logger: Any = None) -> int:
    if ctx.scheduler is None or ctx.dedup is None or ctx.worker is None:
        return 0
    added = 0
    for cat in CATEGORIES:
        for url in load_seeds(category=cat):
            new, again = await _enqueue_seed(ctx, url)
            added += new + again
    return added


async def _handle_crawled(ctx: Any, url: str, result: Any) -> None:
    sender = getattr(ctx, "index_submit_sender", None)
    if sender is not None:                                  # DMZ crawler: hand the page to the private indexers
        acked = await sender.send_to_peers(sender.build_submit_message(result.page, result.discovered_links))
        logger.info("index_submit_sent", url=url, targets=len(sender.submit_peers), acked=acked)
        return
    await _index_and_publish(ctx, result)
    crawl_cfg = ctx.config.crawl
    monitor = getattr(ctx, "feed_monitor", None)
    if crawl_cfg.rss_enabled and crawl_cfg.rss_discovery and monitor is not None:
        for feed_url in result.discovered_feeds:
            if len(monitor.feeds) < crawl_cfg.rss_max_feeds:
                monitor.add_feed(feed_url)


async def seed_and_crawl_loop(ctx: Any, seed_category: str = "tech-docs", *, max_pages: int | None = None) -> int:
    """Runs until cancelled (or ``max_pages`` successful crawls, used by tests and ``infomesh crawl --seeds``).
    Returns the number of pages crawled."""
    if ctx.worker is None or ctx.scheduler is None or ctx.dedup is None:
        logger.warning("seed_and_crawl_loop_skipped", reason="crawler components not initialized (search-only role?)")
        return 0
    seeds = load_seeds(category=seed_category)
    new = again = 0
    for url in seeds:
        n, a = await _enqueue_seed(ctx, url)
        new, again = new + n, again + a
    logger.info("seeds_queued", category=seed_category, total=len(seeds), new=new, rediscovered=again)
    ctx.scheduler.set_urls_per_hour(0)                       # the background loop is not rate-limited per hour
    feed_task = asyncio.create_task(feed_poll_loop(ctx)) if (getattr(ctx, "feed_monitor", None) is not None
                                                             and ctx.config.crawl.rss_enabled) else None
    crawled = 0
    now = time.monotonic()
    last = {"gov": 0.0, "disk": 0.0, "prio": now, "fts": now, "crawl": now}
    try:
        while max_pages is None or crawled < max_pages:
            now = time.monotonic()
            if now - last["gov"] >= GOVERNOR_CHECK_INTERVAL:
                last["gov"] = now
                if await _apply_governor_backpressure(ctx):
                    continue
            if now - last["disk"] > DISK_CHECK_INTERVAL:
                last["disk"] = now
                if is_disk_critically_low(ctx.config.node.data_dir):
                    logger.error("disk_space_critical", msg="Pausing crawl — disk space below 200 MB")
                    await asyncio.sleep(30)
                    continue
            if now - last["prio"] >= PRIORITY_CHECK_INTERVAL:
                last["prio"] = now
                try:
                    n = await _process_priority_queue(ctx)
                    if n:
                        crawled += n
                        last["crawl"] = time.monotonic()
                except Exception:  # noqa: BLE001
                    logger.exception("priority_queue_failed")
            try:
                url, depth = await asyncio.wait_for(ctx.scheduler.get_url(), timeout=5.0)
            except (TimeoutError, asyncio.TimeoutError):
                if time.monotonic() - last["crawl"] >= IDLE_RESTART_SECONDS:
                    try:
                        n = await _reseed_queue(ctx)
                    except Exception:  # noqa: BLE001
                        n = 0
                    if n:
                        last["crawl"] = time.monotonic()
                        logger.info("crawl_reseed_complete", new_urls=n)
                    else:
                        await asyncio.sleep(5)
                else:
                    await asyncio.sleep(1)
                continue
            last["crawl"] = time.monotonic()
            try:
                res = await ctx.worker.crawl_url(url, depth=depth)
                if res.success and res.page:
                    await _handle_crawled(ctx, url, res)
                    crawled += 1
                    _credit(ctx, url[:120])
                elif not res.success:
                    logger.debug("crawl_skipped", url=url, reason=res.error)
            except Exception:  # noqa: BLE001 — one bad page must not end the loop
                logger.exception("crawl_iteration_failed")
            if time.monotonic() - last["fts"] >= FTS_OPTIMIZE_INTERVAL:
                last["fts"] = time.monotonic()
                with contextlib.suppress(Exception):
                    ctx.store.optimize()
    finally:
        if feed_task is not None:
            feed_task.cancel()
            with contextlib.suppress(asyncio.CancelledError):
                await feed_task
    return crawled
