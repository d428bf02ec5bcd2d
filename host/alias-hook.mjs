// alias-hook.mjs — registers the module-resolution hook that maps the reference's '@/...' imports (tsconfig paths,
// e.g. src/modems/fsk.ts:1-3) to '<reference-root>/src/...', so that the UNMODIFIED reference sources load under
//   node --experimental-transform-types --import ./host/alias-hook.mjs host/ref_node_bench.mjs <reference-root>
// The root comes from WAM_REFERENCE_ROOT or the first script argument.  NOT EXECUTED in the build environment (no Node).
import { register } from 'node:module';
import path from 'node:path';
import { pathToFileURL } from 'node:url';

const root = path.resolve(process.env.WAM_REFERENCE_ROOT ?? process.argv[2] ?? '.');
register('./alias-hook-impl.mjs', { parentURL: import.meta.url, data: { srcURL: pathToFileURL(path.join(root, 'src') + path.sep).href } });
