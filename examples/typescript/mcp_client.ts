// Talk to a node over MCP.  Either spawn it over stdio (default) or connect to the streamable-HTTP endpoint:
//   npx tsx mcp_client.ts "tensor memory"                       -> spawns `infomesh mcp`
//   INFOMESH_MCP_URL=http://127.0.0.1:8081/mcp npx tsx mcp_client.ts "tensor memory"
import { Client } from "@modelcontextprotocol/sdk/client/index.js";
import { StdioClientTransport } from "@modelcontextprotocol/sdk/client/stdio.js";
import { StreamableHTTPClientTransport } from "@modelcontextprotocol/sdk/client/streamableHttp.js";

const query = process.argv[2] ?? "python asyncio tutorial";
const url = process.env.INFOMESH_MCP_URL;
const apiKey = process.env.INFOMESH_API_KEY;           // only needed when the node was started with one

const transport = url
  ? new StreamableHTTPClientTransport(new URL(url))
  : new StdioClientTransport({ command: "infomesh", args: ["mcp"] });
const client = new Client({ name: "infomesh-ts-example", version: "0.1.0" });
await client.connect(transport);

const { tools } = await client.listTools();
console.log("tools:", tools.map((t) => t.name).join(", "));

const text = (r: any): string => r.content?.map((c: any) => c.text ?? "").join("\n") ?? "";
const call = (name: string, args: Record<string, unknown>) =>
  client.callTool({ name, arguments: apiKey ? { ...args, api_key: apiKey } : args });

// 1. search (JSON so the result can be consumed programmatically); on a GPU node this is the hybrid pipeline
const found = JSON.parse(text(await call("web_search", { query, top_k: 5, format: "json" })));
for (const r of found.results ?? []) console.log(`${r.score?.toFixed?.(3) ?? "-"}  ${r.title}  ${r.url}`);

// 2. fetch the full text of the best hit (served from the index; crawled on a miss)
if (found.results?.length) {
  const page = text(await call("fetch_page", { url: found.results[0].url }));
  console.log("\n--- first 400 characters ---\n" + page.slice(0, 400));
}

// 3. node status: index size, peers, credits, GPU plane
console.log("\n" + text(await call("status", {})));
await client.close();
