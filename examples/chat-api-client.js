// Node.js client for `dllama-api` (same role as the reference's examples/chat-api-client.js).
// Usage: ./dllama-api --model m.m --tokenizer t.t --port 9990 &   then   node examples/chat-api-client.js
const HOST = process.env.HOST ? process.env.HOST : '127.0.0.1';
const PORT = process.env.PORT ? Number(process.env.PORT) : 9990;

async function complete(messages, maxTokens) {
    const response = await fetch(`http://${HOST}:${PORT}/v1/chat/completions`, {
        method: 'POST',
        headers: { 'Content-Type': 'application/json' },
        body: JSON.stringify({ messages, temperature: 0.7, stop: ['<|eot_id|>'], max_tokens: maxTokens }),
    });
    return await response.json();
}

(async () => {
    const history = [{ role: 'system', content: 'You are an excellent math teacher.' }];
    for (const question of ['What is 1 + 2?', 'And multiplied by 4?']) {
        history.push({ role: 'user', content: question });
        const answer = await complete(history, 128);          // the server reuses the KV prefix of earlier turns
        const text = answer.choices[0].message.content;
        history.push({ role: 'assistant', content: text });
        console.log(`> ${question}\n${text}\n`, answer.usage);
    }
})();
