"""A tiny R lexer for the CPU suite (there is no R in the build image): strips comments and string
literals so that delimiter balance and identifier use of r/ccgp.R and tools/r_crosscheck.R can be checked."""


def strip_strings_and_comments(src: str) -> str:
    out = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c == "#":
            while i < n and src[i] != "\n":
                i += 1
            continue
        if c in "\"'`":
            q = c
            i += 1
            while i < n and src[i] != q:
                i += 2 if src[i] == "\\" else 1
            i += 1
            out.append('""' if q != "`" else "bq")
            continue
        out.append(c)
        i += 1
    return "".join(out)


def check_balanced(src: str):
    """-> None when (), [], {} nest properly; else (line, message)."""
    code = strip_strings_and_comments(src)
    stack = []
    line = 1
    pairs = {")": "(", "]": "[", "}": "{"}
    for ch in code:
        if ch == "\n":
            line += 1
        elif ch in "([{":
            stack.append((ch, line))
        elif ch in ")]}":
            if not stack or stack[-1][0] != pairs[ch]:
                return line, "unexpected %r" % ch
            stack.pop()
    if stack:
        return stack[-1][1], "unclosed %r" % stack[-1][0]
    return None
