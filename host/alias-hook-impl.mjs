// alias-hook-impl.mjs — the hooks themselves (run on Node's hooks thread; registered by alias-hook.mjs).
// '@/x/y' -> '<root>/src/x/y.ts' (or '/index.ts'); relative imports without an extension get '.ts' as the reference's
// bundler would resolve them.  Everything else goes to the default resolver.
import { existsSync } from 'node:fs';
import { fileURLToPath } from 'node:url';

let srcURL;
export function initialize(data) {
  srcURL = data.srcURL;
}

function withTsExtension(url) {
  const file = fileURLToPath(url);
  for (const candidate of [file, `${file}.ts`, `${file}/index.ts`]) {
    if (existsSync(candidate) && !candidate.endsWith('/')) return new URL(candidate === file ? url : candidate === `${file}.ts` ? `${url}.ts` : `${url}/index.ts`).href;
  }
  return url;
}

export async function resolve(specifier, context, nextResolve) {
  if (specifier.startsWith('@/')) return nextResolve(withTsExtension(new URL(specifier.slice(2), srcURL).href), context);
  if ((specifier.startsWith('./') || specifier.startsWith('../')) && context.parentURL?.startsWith(srcURL) && !/\.[cm]?[jt]s$/.test(specifier))
    return nextResolve(withTsExtension(new URL(specifier, context.parentURL).href), context);
  return nextResolve(specifier, context);
}
