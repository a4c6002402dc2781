// Plain HTTP against the local admin API (no SDK needed): health, search, index stats, Prometheus metrics.
//   INFOMESH_ADMIN_URL=http://127.0.0.1:8080 INFOMESH_API_KEY=... npx tsx http_client.ts "thread block clusters"
const base = process.env.INFOMESH_ADMIN_URL ?? "http://127.0.0.1:8080";
const headers: Record<string, string> = process.env.INFOMESH_API_KEY ? { "x-api-key": process.env.INFOMESH_API_KEY } : {};

async function get<T>(path: string, accept = "application/json"): Promise<T> {
  const res = await fetch(base + path, { headers: { ...headers, accept } });
  if (res.status === 429) throw new Error("rate limited: the admin API allows 10 requests per second per client");
  if (!res.ok) throw new Error(`${path}: HTTP ${res.status}`);
  return (accept === "application/json" ? res.json() : res.text()) as Promise<T>;
}

type Hit = { url: string; title: string; snippet: string; score: number };
const q = process.argv[2] ?? "python asyncio tutorial";

console.log("health:", await get<Record<string, string>>("/health?detail=1"));
const out = await get<{ total: number; elapsed_ms: number; results: Hit[] }>(`/search?q=${encodeURIComponent(q)}&limit=5`);
console.log(`${out.total} results in ${out.elapsed_ms} ms`);
for (const h of out.results) console.log(`  ${h.score.toFixed(3)}  ${h.title}\n           ${h.url}`);
console.log("index:", await get("/index/stats"));
console.log("gpu:", await get("/gpu/stats"));
const metrics = await get<string>("/metrics", "text/plain");
console.log(metrics.split("\n").filter((l) => l.startsWith("infomesh_search")).join("\n"));
